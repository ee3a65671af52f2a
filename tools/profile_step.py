#!/usr/bin/env python
"""One-GPU driver for ncu captures: N resident steps of the BASELINE config-2 workload (64 VGA pairs, top_k 4096).
    ncu --set full --clock-control none --import-source on -k regex:mnn_tc_kernel -s 2 -c 1 -o gpurun_out/prof_mnn \
        python tools/profile_step.py --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from accelerated_features_b200 import XFeat  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--workload", default="sparse", choices=["sparse", "star"])
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--width", type=int, default=640)
a = ap.parse_args()
xf = XFeat(top_k=4096)
g = torch.Generator().manual_seed(0)
x1 = torch.randn(a.batch, 3, a.height, a.width, generator=g).cuda()
x2 = torch.randn(a.batch, 3, a.height, a.width, generator=g).cuda()
import time
for i in range(a.steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if a.workload == "sparse":
        mk0, mk1, cnt = xf._match_sparse_batch_device(x1, x2, 4096, -1)
    else:
        m, n = xf._match_star_device(x1, x2, 4096)
    torch.cuda.synchronize()
    print(f"step {i}: {1e3 * (time.perf_counter() - t0):.2f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
if a.workload == "star":
    print("refined matches per pair (first 8):", n[:8].tolist())
print("done", a.workload, a.steps)
