#!/usr/bin/env python
"""Secondary BASELINE.json configs (not bench.py lines): C1 single-pair latency and the C5 MNN sweep 2k..32k (C3, the semi-dense
1280x960 batch, is `bench.py --config star`).  Writes one JSON document to stdout.
    python tools/bench_configs.py > gpurun_out/configs.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from accelerated_features_b200 import XFeat, weights as W  # noqa: E402

xf = XFeat(weights=W.load_state_dict(W.DEFAULT_WEIGHTS))
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
out = {}


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# C1: one VGA pair through the public API (numpy uint8 in, numpy out), wall clock incl. H2D/D2H
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "inputs_assets_vga.npz"))
ref, tgt = g["ref"], g["tgt"]
for _ in range(3):
    mk0, mk1 = xf.match_xfeat(ref, tgt, top_k=4096)
t0 = time.perf_counter()
for _ in range(20):
    mk0, mk1 = xf.match_xfeat(ref, tgt, top_k=4096)
out["C1_single_vga_pair"] = {"ms_per_pair_wall": 1e3 * (time.perf_counter() - t0) / 20, "matches": int(len(mk0)),
                             "note": "asset pair resized to 640x480, numpy uint8 in -> numpy out, includes H2D/D2H and host sync"}

# C5: MNN sweep, unit-norm 64-D: default (1: three-term split, persistent), 4 (fp16 filter + exact re-score), 0 (fp32 CUDA cores)
sweep = []
for n in (2048, 4096, 8192, 16384, 32768):
    gen = torch.Generator().manual_seed(n)
    f1 = F.normalize(torch.randn(n, 64, generator=gen), dim=-1).cuda()
    f2 = F.normalize(torch.randn(n, 64, generator=gen), dim=-1).cuda()
    row = {"n": n}
    for impl, name in ((1, "tcgen05_three_term_default"), (4, "tcgen05_filter_plus_exact"), (0, "fp32_simt")):
        xf._lib.xfeat_set_mnn_impl(impl)
        ms = timed(lambda: xf._mnn_device(f1, None, n, 0, f2, None, n, 0, 1, -1), n=10, warm=3)
        flops = 2.0 * n * n * 64
        alg_bytes = 2 * n * (256 + 8)
        ref_bytes = alg_bytes + 2.0 * n * n * 4
        row[name] = {"ms": ms, "algorithmic_tflops": flops / ms / 1e9, "frac_of_bf16_sustained": flops / ms / 1e9 / peaks["bf16_tflops_sustained"],
                     "algorithmic_gbs": alg_bytes / ms / 1e6, "reference_equivalent_gbs": ref_bytes / ms / 1e6}
    idx0, idx1, cnt = xf._mnn_device(f1, None, n, 0, f2, None, n, 0, 1, -1)
    row["mutual_matches"] = int(cnt.item())
    sweep.append(row)
xf._lib.xfeat_set_mnn_impl(1)
out["C5_mnn_sweep"] = {"rows": sweep, "note": "single pair per call (a 32k x 32k pair fills the GPU; small N under-fills 148 SMs); "
                       "reference-equivalent GB/s = bytes the reference's materialised S (written + read) would move / our time; "
                       f"HBM copy peak {peaks['hbm_gbs']} GB/s"}
print(json.dumps(out, indent=1))
