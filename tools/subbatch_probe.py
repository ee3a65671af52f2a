#!/usr/bin/env python
"""Does running the backbone in L2-sized sub-batches pay?  Times xfeat_net on 128 VGA images as one launch sequence and as
128/sb sequences of sb images that reuse the same (then L2-resident) activation workspace."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from accelerated_features_b200 import XFeat  # noqa: E402

xf = XFeat(top_k=4096)
B, H, W = 128, 480, 640
g = torch.Generator().manual_seed(0)
xn = torch.randn(B, H, W, generator=g).cuda()
res = {}
for sb in (128, 64, 32, 16, 8):
    def run():
        for o in range(0, B, sb):
            xf._run_net(xn[o:o + sb], sb, H, W)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    res[f"sub_batch_{sb}_ms"] = e0.elapsed_time(e1) / 10
print(json.dumps(res))
