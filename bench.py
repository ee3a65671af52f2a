#!/usr/bin/env python
"""bench.py -- image-pairs/sec (extract + match) on synthetic VGA batches (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of 64 synthetic VGA image pairs per GPU:
    detectAndCompute on both image sets (backbone, NMS/top-k 4096, bicubic descriptors) + per-pair MNN match.
`value`  : pairs/s with the inputs already resident in HBM (CUDA events around K steps, max over ranks).
`e2e`    : pairs/s through the public batch API path with HOST (pinned) inputs: H2D of both image sets and D2H of the
           matched keypoints + counts inside the timed region, double-buffered on a copy stream.
`roofline`: the dominant kernel (MNN D1.D2^T + fused arg-max): algorithmic FLOPs / its CUDA-event time inside the timed steps.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-pairs/sec (extract+match) VGA batch=64"
H, W, BATCH, TOPK = 480, 640, 64, 4096
WORKLOAD = "batch=64 synthetic VGA (640x480) sparse match_xfeat top_k=4096"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the upper half = samples taken under load
        sm_sorted = sorted(sm)
        return {"sm_mhz": statistics.median(sm_sorted[len(sm_sorted) // 2:]), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port, torch CPU = the arithmetic the reference itself uses on CPU)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_pairs_per_s(n_pairs: int, repeats: int, warmup: int, seed: int = 0):
    import torch
    from oracle import xfeat_oracle as orc
    # all the host threads torch would use by default (physical cores); torchrun exports OMP_NUM_THREADS=1, undo that here
    want = max(1, (os.cpu_count() or 2) // 2)
    if torch.get_num_threads() < want:
        torch.set_num_threads(want)
    sd = orc.load_state()
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(n_pairs, 3, H, W, generator=g)
    x2 = torch.randn(n_pairs, 3, H, W, generator=g)

    def step():
        with torch.inference_mode():
            o1 = orc.detect_and_compute(sd, x1, TOPK)
            o2 = orc.detect_and_compute(sd, x2, TOPK)
            n = 0
            for a, b in zip(o1, o2):
                i0, i1 = orc.mnn_match(a["descriptors"], b["descriptors"], -1)
                n += len(i0)
        return n

    for _ in range(warmup):
        step()
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return n_pairs / statistics.median(times), sum(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n_pairs = 8
    v, total_s, threads = cpu_pairs_per_s(n_pairs, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_pairs / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic randn",
        "config": {"workload": WORKLOAD, "sample": f"{n_pairs} pairs per step on the host CPU", "top_k": TOPK},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": f"{n_pairs} VGA pairs/step, oracle port of the reference algorithm (torch {torch.__version__} CPU, "
                                   f"{threads} threads of {os.cpu_count()} logical cores)"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from accelerated_features_b200 import XFeat, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the XFeat hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    from accelerated_features_b200 import weights as _w
    xf = XFeat(weights=_w.load_state_dict(_w.DEFAULT_WEIGHTS), top_k=TOPK, device=local)   # dict: no 'loading weights' print on stdout
    lib = _lib.load()

    # synthetic data: the reference's own style (minimal_example.py: torch.randn), one distinct shard per rank
    g = torch.Generator().manual_seed(1000 + rank)
    h1 = torch.randn(BATCH, 3, H, W, generator=g).pin_memory()
    h2 = torch.randn(BATCH, 3, H, W, generator=g).pin_memory()
    d1, d2 = h1.to(dev), h2.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mnn_events = []

    def step_resident(record=False):
        o = xf._detect_sparse_device([d1, d2], TOPK, xf.detection_threshold)
        k1, k2 = o["keypoints"][:BATCH], o["keypoints"][BATCH:]
        f1, f2 = o["descriptors"][:BATCH], o["descriptors"][BATCH:]
        n1, n2 = o["n_valid"][:BATCH], o["n_valid"][BATCH:]
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        idx0, idx1, cnt = xf._mnn_device(f1, n1, TOPK, TOPK * 64, f2, n2, TOPK, TOPK * 64, BATCH, -1)
        if record:
            e1.record()
            mnn_events.append((e0, e1))
        mk0, mk1 = xf._empty((BATCH, TOPK, 2)), xf._empty((BATCH, TOPK, 2))
        _lib.check(lib.xfeat_gather_matches(k1.data_ptr(), k2.data_ptr(), TOPK, TOPK, idx0.data_ptr(), idx1.data_ptr(),
                                            cnt.data_ptr(), BATCH, mk0.data_ptr(), mk1.data_ptr(), xf._stream()), "gather")
        return mk0, mk1, cnt, n1, n2

    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    barrier()
    launches0 = lib.xfeat_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        out = step_resident(record=True)
    t1.record()
    barrier()
    launches = lib.xfeat_launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    mnn_ms = statistics.mean(a.elapsed_time(b) for a, b in mnn_events)
    n1_mean = float(out[3].float().mean())
    n2_mean = float(out[4].float().mean())
    matches_mean = float(out[2].float().mean())

    # ---- end to end: pinned host inputs -> H2D (copy stream, double buffered) -> kernels -> D2H of results ----
    copy_stream = torch.cuda.Stream(dev)
    bufs = [(torch.empty_like(d1), torch.empty_like(d2)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    r0 = torch.empty((BATCH, TOPK, 2), dtype=torch.float32).pin_memory()
    r1 = torch.empty((BATCH, TOPK, 2), dtype=torch.float32).pin_memory()
    rc = torch.empty((BATCH,), dtype=torch.int32).pin_memory()
    main = torch.cuda.current_stream(dev)

    # uint8 HWC images (what cv2 / camera callers hand to match_xfeat): 4x fewer bytes over PCIe, "/255" on the device
    hu1 = (torch.rand(BATCH, H, W, 3, generator=g) * 255).to(torch.uint8).pin_memory()
    hu2 = (torch.rand(BATCH, H, W, 3, generator=g) * 255).to(torch.uint8).pin_memory()
    ubufs = [(torch.empty_like(hu1, device=dev), torch.empty_like(hu2, device=dev)) for _ in range(2)]

    def e2e_run(n, u8=False):
        src1, src2, bb = (hu1, hu2, ubufs) if u8 else (h1, h2, bufs)

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i % 2])
                bb[i % 2][0].copy_(src1, non_blocking=True)
                bb[i % 2][1].copy_(src2, non_blocking=True)
                ready[i % 2].record(copy_stream)

        for ev in freed:
            ev.record(main)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            main.wait_event(ready[i % 2])
            if u8:
                mk0, mk1, cnt = xf._match_sparse_batch_device(bb[i % 2][0].permute(0, 3, 1, 2), bb[i % 2][1].permute(0, 3, 1, 2),
                                                              TOPK, -1, div255=True)
            else:
                mk0, mk1, cnt = xf._match_sparse_batch_device(bb[i % 2][0], bb[i % 2][1], TOPK, -1)
            freed[i % 2].record(main)
            r0.copy_(mk0, non_blocking=True)
            r1.copy_(mk1, non_blocking=True)
            rc.copy_(cnt, non_blocking=True)

    e2e_run(2)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_run(args.steps)
    s1.record()
    barrier()
    e2e_ms = s0.elapsed_time(s1)
    e2e_run(2, u8=True)
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    e2e_run(args.steps, u8=True)
    u1.record()
    barrier()
    e2e_u8_ms = u0.elapsed_time(u1)
    # host link alone: the same fp32 upload with no kernels behind it
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        bufs[0][0].copy_(h1, non_blocking=True)
        bufs[0][1].copy_(h2, non_blocking=True)
    c1.record()
    barrier()
    h2d_gbs = 3 * 2 * h1.numel() * 4 / (c0.elapsed_time(c1) / 1e3) / 1e9
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([ms_total, e2e_ms, mnn_ms, e2e_u8_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, mnn_ms, e2e_u8_ms = t.tolist()

    if rank == 0:
        pk, pk_kind = peaks()
        pairs = BATCH * world * args.steps
        value = pairs / (ms_total / 1e3)
        e2e = pairs / (e2e_ms / 1e3)
        # dominant kernel: mnn_scan_kernel; algorithmic FLOPs = 2 * n1 * n2 * 64 per pair, summed over the batch
        flops = 2.0 * 64.0 * float((out[3].double() * out[4].double()).sum())
        achieved = flops / (mnn_ms / 1e3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        tc = lib.xfeat_get_mnn_impl() >= 1
        impl = lib.xfeat_get_mnn_impl()
        kname = {1: "mnn_tc_kernel (tcgen05 split-fp16, 3 K=64 blocks D1.D2^T, fp32 accumulate in TMEM, fused row arg-max, both directions",
                 2: "mnn_tc_once_kernel (tcgen05 split-fp16, one GEMM, row + column arg-max in the epilogue",
                 3: "mnn_tc2_kernel (tcgen05 cta_group::2 CTA pairs, split-fp16, fused row arg-max, both directions"}.get(
            impl, "mnn_scan_kernel (fp32 FFMA D1.D2^T + fused row/col arg-max") + "; timed call also contains split/finalize)"
        ncu = None
        for name in ("ncu_mnn_tc_once.json", "ncu_mnn_tc_final.json"):
            try:
                with open(os.path.join(ROOT, "profiles", "r01", name)) as f:
                    ncu = json.load(f)
                if (ncu.get("kernel") == "mnn_tc_once_kernel") == (impl == 2):
                    break
                ncu = None
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic randn (reference minimal_example.py style), pretrained XFeat weights",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": BATCH, "top_k": TOPK, "l2": "inputs 472 MB/step > 126 MB L2",
                       "mean_keypoints": [n1_mean, n2_mean], "mean_matches_per_pair": matches_mean, "parallelism": f"pair-sharded x{world}"},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * h1.numel() * 4),
                    "d2h_bytes_per_step": int(2 * r0.numel() * 4 + rc.numel() * 4), "ms_per_step": e2e_ms / args.steps,
                    "h2d_only_gbs": h2d_gbs,
                    "note": "fp32 images (the reference's synthetic input style): bound by the host link, see h2d_only_gbs"},
            "e2e_u8": {"value": pairs / (e2e_u8_ms / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(2 * hu1.numel()),
                       "ms_per_step": e2e_u8_ms / args.steps, "input": "uint8 HWC images, /255 on device (camera / cv2 callers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": kname, "bound": "tensor",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_kind": f"{pk_kind} bf16 dense, sustained", "ms_per_launch": mnn_ms,
                         "note": "achieved counts each MAC once (SURVEY 8d); the kernel executes 3 fp16 split passes per MAC for fp32-equivalent "
                                 "results, so tensor-pipe work is 3x this figure",
                         "tensor_work_frac": 3.0 * achieved / peak if tc else None,
                         "traffic": (ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) if (ncu and tc) else None,
                         "ncu": ncu if tc else None},
        }
        if world == 1 and not args.no_cpu:
            v, total_s, threads = cpu_pairs_per_s(8, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                                    "sample": f"3 x 8 VGA pairs, oracle port (torch CPU, {threads} threads of {os.cpu_count()} logical cores), {total_s:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
