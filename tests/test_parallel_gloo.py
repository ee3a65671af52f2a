"""Host-side multi-GPU logic on CPU with the gloo backend (world_size 2 and 3): pair sharding and the image-sharded
semi-dense layout return exactly what a single process returns."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from accelerated_features_b200 import parallel as par


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [par.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


# deterministic stand-ins for the GPU stages (per-image / per-pair functions of the data only)
def fake_extract(imgs):
    n = imgs.shape[0]
    base = imgs.reshape(n, -1)[:, :6]
    return {"keypoints": base[:, None, :2].repeat(1, 5, 1) + torch.arange(5.)[None, :, None],
            "descriptors": (base[:, None, :] * torch.arange(1, 6.)[None, :, None]).repeat(1, 1, 11)[:, :, :64].contiguous(),
            "scales": base[:, None, 0].repeat(1, 5)}


def fake_match_refine(d1, d2):
    out = []
    for b in range(d1["descriptors"].shape[0]):
        s = d1["descriptors"][b] @ d2["descriptors"][b].t()
        out.append((s.argmax(1) + 1000 * (d1["keypoints"][b, :, 0] + d2["scales"][b]).long()).numpy())
    return out


def fake_match(i1, i2):
    return [float((a * b).sum()) for a, b in zip(i1, i2)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        s1 = torch.randn(6, 3, 8, 8, generator=g)
        s2 = torch.randn(6, 3, 8, 8, generator=g)
        star = par.star_image_sharded(fake_extract, fake_match_refine, s1, s2)
        pairs = par.match_pairs_sharded(fake_match, s1[:5], s2[:5])      # 5 pairs: ragged split
        blk = par.all_gather_blocks(torch.full((2, 3), float(rank)))
        q.put((rank, [x.tolist() for x in star], pairs, blk.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_layouts_match_single_process(world):
    g = torch.Generator().manual_seed(0)
    s1 = torch.randn(6, 3, 8, 8, generator=g)
    s2 = torch.randn(6, 3, 8, 8, generator=g)
    want_star = [x.tolist() for x in fake_match_refine(fake_extract(s1), fake_extract(s2))]
    want_pairs = fake_match(s1[:5], s2[:5])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, star, pairs, blk in got:
        assert star == want_star                 # byte-identical to the single-process result, on every rank
        assert pairs == want_pairs
        assert blk == [[float(r)] * 3 for r in range(world) for _ in range(2)]
