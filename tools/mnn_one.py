#!/usr/bin/env python
"""ncu driver: `python tools/mnn_one.py <impl> [calls] [N]` -- xfeat_mnn_match on 64 pairs of N (default 4096) unit-norm randn
descriptors with one implementation, for `ncu -k regex:mnn -s <skip> -c 1 --set full ...` captures."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from accelerated_features_b200 import XFeat  # noqa: E402

impl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
B = 64
xf = XFeat(top_k=4096)
g = torch.Generator().manual_seed(0)
f1 = F.normalize(torch.randn(B, N, 64, generator=g), dim=-1).cuda()
f2 = F.normalize(torch.randn(B, N, 64, generator=g), dim=-1).cuda()
xf._lib.xfeat_set_mnn_impl(impl)
for _ in range(calls):
    r = xf._mnn_device(f1, None, N, N * 64, f2, None, N, N * 64, B, -1, abs_bound=1.0)
torch.cuda.synchronize()
print("matches", r[2][:4].tolist())
