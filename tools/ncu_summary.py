#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion needs."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_op",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg", "smsp__inst_executed.sum ",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fma_cycles_active.avg.pct", "smsp__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct", "sm__maximum_warps_per_active_cycle_pct", "launch__occupancy_limit",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        lines.append(f"## {name[:140]}  grid {vals[hdr.index('Grid Size')]} block {vals[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS):
                lines.append(f"- `{h}` = {vals[i]} {units[i]}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "a").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
