#!/usr/bin/env python
"""Run under torchrun on N >= 2 GPUs: the pair-sharded sparse path and the image-sharded semi-dense path must return
byte-identical results to a single-GPU run of the same inputs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from accelerated_features_b200 import XFeat, parallel as par  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
xf = XFeat(device=local)
g = torch.Generator().manual_seed(7)
B = 2 * world
s1 = torch.randn(B, 3, 256, 320, generator=g)
s2 = torch.randn(B, 3, 256, 320, generator=g)

# ---- sparse, pair-sharded: no data-path collective ----
sharded = par.match_pairs_sharded(lambda a, b: xf.match_xfeat_batch(a, b, top_k=1024), s1, s2)
ok = True
if rank == 0:
    single = xf.match_xfeat_batch(s1, s2, top_k=1024)
    ok &= len(sharded) == B and all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(sharded, single))
    print(f"[rank0] sparse pair-sharded x{world}: identical to single GPU = {ok}; matches per pair {[len(a[0]) for a in single]}")


# ---- semi-dense, image-sharded: one NCCL all-gather of the coarse feature blocks ----
def extract(imgs):
    return xf._dense_device(imgs, 2048, True)


def match_refine(d1, d2):
    K = d1["descriptors"].shape[1]
    d1 = {k: v.contiguous() for k, v in d1.items()}
    d2 = {k: v.contiguous() for k, v in d2.items()}
    idx0, idx1, cnt = xf._mnn_device(d1["descriptors"], None, K, K * 64, d2["descriptors"], None, K, K * 64, len(cnt_dummy(d1)), -1)
    m, n = xf._refine_device(d1, d2, idx0, idx1, cnt)
    return [m[b, :int(n[b])].cpu().numpy() for b in range(m.shape[0])]


def cnt_dummy(d):
    return range(d["descriptors"].shape[0])


star = par.star_image_sharded(extract, match_refine, s1, s2)
# time it once more (warm) and say what crossed NVLink: every rank all-gathers the coarse blocks of its 2B/world images
torch.cuda.synchronize(); dist.barrier()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
par.star_image_sharded(extract, match_refine, s1, s2)
t1.record()
torch.cuda.synchronize()
ms = torch.tensor([t0.elapsed_time(t1)], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    kk = int(2048 * 0.2) + int(2048 * 0.8)
    bytes_per_image = kk * (64 + 2 + 1) * 4
    recv = (2 * B - 2 * B // world) * bytes_per_image
    print(f"[rank0] star image-sharded x{world}: {float(ms):.2f} ms (max over ranks); all-gather of {2 * B} coarse blocks x {bytes_per_image / 1e6:.2f} MB "
          f"= {recv / 1e6:.2f} MB received per rank over NVLink")
if rank == 0:
    ref = xf.match_xfeat_star(s1, s2, top_k=2048)
    same = len(star) == B and all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(star, ref))
    print(f"[rank0] star image-sharded x{world}: identical to single GPU = {same}; refined per pair {[len(a) for a in star]}")
    ok &= same
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
