#!/usr/bin/env python
"""Probe of the MNN implementations on (a) unit-norm randn descriptors (BASELINE config 5 style) and (b) the descriptors the
sparse path extracts from the bench images: ms per call for implementations 1 and 4, and the fraction of rows the filter
pass of implementation 4 must hand to the exact kernel (top-1/top-2 gap of S~ = hi.hi^T within tau; recomputed with torch)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from accelerated_features_b200 import XFeat  # noqa: E402

xf = XFeat(top_k=4096)
B, N = 64, 4096


def amb_fraction(f1, f2):
    """per pair: fraction of rows (both directions) with gap <= tau, same formula as mnn_fast_kernel (scale-free)."""
    out = []
    for b in range(min(4, f1.shape[0])):
        a, c = f1[b].float(), f2[b].float()
        ah, ch = a.half().float(), c.half().float()
        al, cl = (a - ah).half().float(), (c - ch).half().float()
        s = ah @ ch.t()
        for (S, nh, nl, Hm, Lm) in ((s, ah.norm(dim=1), al.norm(dim=1), ch.norm(dim=1).max(), cl.norm(dim=1).max()),
                                    (s.t(), ch.norm(dim=1), cl.norm(dim=1), ah.norm(dim=1).max(), al.norm(dim=1).max())):
            t2 = torch.topk(S, 2, dim=1).values
            tau = 2.1 * (nh * Lm + nl * Hm) + 6.2e-5 * nh * Hm
            out.append(float(((t2[:, 0] - t2[:, 1]) <= tau).float().mean()))
    return sum(out) / len(out)


def time_impl(impl, f1, n1, f2, n2, bound):
    xf._lib.xfeat_set_mnn_impl(impl)
    for _ in range(3):
        r = xf._mnn_device(f1, n1, N, N * 64, f2, n2, N, N * 64, B, -1, abs_bound=bound)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r = xf._mnn_device(f1, n1, N, N * 64, f2, n2, N, N * 64, B, -1, abs_bound=bound)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10, r


res = {}
g = torch.Generator().manual_seed(0)
f1 = F.normalize(torch.randn(B, N, 64, generator=g), dim=-1).cuda()
f2 = F.normalize(torch.randn(B, N, 64, generator=g), dim=-1).cuda()
res["unit_randn"] = {"amb_fraction": amb_fraction(f1, f2)}
for impl in (1, 4):
    ms, r = time_impl(impl, f1, None, f2, None, 1.0)
    res["unit_randn"][f"impl{impl}_ms"] = ms
x1 = torch.randn(B, 3, 480, 640, generator=g).cuda()
x2 = torch.randn(B, 3, 480, 640, generator=g).cuda()
o = xf._detect_sparse_device([x1, x2], 4096, 0.05)
d1, d2 = o["descriptors"][:B].contiguous(), o["descriptors"][B:].contiguous()
n1, n2 = o["n_valid"][:B].contiguous(), o["n_valid"][B:].contiguous()
res["bench_descriptors"] = {"amb_fraction": amb_fraction(d1, d2)}
for impl in (1, 4):
    ms, r = time_impl(impl, d1, n1, d2, n2, 1.0)
    res["bench_descriptors"][f"impl{impl}_ms"] = ms
xf._lib.xfeat_set_mnn_impl(4)
print(json.dumps(res))
