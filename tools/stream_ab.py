#!/usr/bin/env python
"""match_xfeat_stream end to end on pinned uint8 HWC batches (64 VGA pairs): ms per batch over N batches.
    python tools/stream_ab.py [N]            (XFEAT_STREAM_D2H_MAIN=1 puts the result copies back on the compute stream)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from accelerated_features_b200 import XFeat  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
xf = XFeat(top_k=4096)
g = torch.Generator().manual_seed(0)
a = xf.pinned_like((64, 480, 640, 3)); b = xf.pinned_like((64, 480, 640, 3))
a.copy_((torch.rand(64, 480, 640, 3, generator=g) * 255).to(torch.uint8)); b.copy_((torch.rand(64, 480, 640, 3, generator=g) * 255).to(torch.uint8))
na, nb = a.numpy(), b.numpy()
for _ in xf.match_xfeat_stream(((na, nb) for _ in range(3))):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 0
for res in xf.match_xfeat_stream(((na, nb) for _ in range(N))):
    n += len(res)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"batches": N, "ms_per_batch": 1e3 * dt / N, "pairs_per_s": n / dt, "d2h_on_main": bool(os.environ.get("XFEAT_STREAM_D2H_MAIN"))}))
