"""Multi-GPU orchestration for the XFeat hot path: one process per GPU, `torch.distributed` for the plumbing.

The path shards by independent units: an image for extraction, an image *pair* for matching.
  * pair-sharded (default): rank r owns a contiguous slice of the pairs, both images of a pair live on the same GPU
    -> no data-path collective at all; results are gathered once at the end (fixed-size blocks or Python objects).
  * image-sharded semi-dense (`star_image_sharded`): the 2B images are spread over the ranks for load balance, so the
    two images of a pair may sit on different GPUs -> ONE all-gather of the coarse descriptor / keypoint / scale blocks
    (NCCL over NVLink), then every rank matches + refines its pair slice.  Returns byte-identical results to the
    pair-sharded layout because the per-image extraction and the per-pair match do not depend on the batch composition.
Everything here is host logic over callables, so it is exercised on CPU with the gloo backend (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of range(n_items); the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_blocks(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate equally shaped per-rank blocks along dim 0 (one collective)."""
    world, _ = _world(group)
    if world == 1:
        return t
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, t, group=group)
    else:
        dist.all_gather(list(out.chunk(world, 0)), t, group=group)
    return out


def gather_objects(local: Sequence, group=None) -> List:
    """Gather per-rank Python result lists (variable-length matches) in rank order on every rank."""
    world, _ = _world(group)
    if world == 1:
        return list(local)
    buckets: List = [None] * world
    dist.all_gather_object(buckets, list(local), group=group)
    return [x for b in buckets for x in b]


def match_pairs_sharded(match_fn: Callable, imgs1, imgs2, group=None) -> List:
    """Pair-sharded batch matching: `match_fn(imgs1[a:b], imgs2[a:b]) -> list` runs on this rank's slice, results of all
    pairs are returned on every rank in pair order. No collective touches the data path."""
    world, rank = _world(group)
    a, b = shard_range(len(imgs1), world, rank)
    local = match_fn(imgs1[a:b], imgs2[a:b]) if b > a else []
    return gather_objects(local, group)


def star_image_sharded(extract_fn: Callable[[torch.Tensor], Dict[str, torch.Tensor]],
                       match_refine_fn: Callable[[Dict[str, torch.Tensor], Dict[str, torch.Tensor]], List],
                       im_set1: torch.Tensor, im_set2: torch.Tensor, group=None) -> List:
    """Image-sharded semi-dense matching.

    extract_fn(images (n,C,H,W)) -> {'keypoints' (n,K,2), 'descriptors' (n,K,64), 'scales' (n,K)}   (detectAndComputeDense)
    match_refine_fn(d1, d2) -> list of per-pair results for the given coarse feature batches          (batch_match + refine)
    The 2B images (set 1 then set 2) are split evenly; 2B must be divisible by the world size so that the all-gather
    moves equal blocks."""
    world, rank = _world(group)
    B = im_set1.shape[0]
    if im_set2.shape != im_set1.shape:
        raise RuntimeError("both image sets must have the same shape")
    if (2 * B) % world:
        raise RuntimeError("image-sharded layout needs 2*B divisible by the world size")
    per = 2 * B // world
    lo, hi = rank * per, (rank + 1) * per
    # this rank's images, taken from the virtual concatenation [set1 ; set2] without materialising it
    parts = []
    if lo < B:
        parts.append(im_set1[lo:min(hi, B)])
    if hi > B:
        parts.append(im_set2[max(lo, B) - B:hi - B])
    feats = [extract_fn(p) for p in parts]
    local = {k: torch.cat([f[k] for f in feats], 0) if len(feats) > 1 else feats[0][k] for k in feats[0]}
    full = {k: all_gather_blocks(v, group) for k, v in local.items()}          # the one exchange step
    a, b = shard_range(B, world, rank)
    d1 = {k: v[a:b] for k, v in full.items()}
    d2 = {k: v[B + a:B + b] for k, v in full.items()}
    res = match_refine_fn(d1, d2) if b > a else []
    return gather_objects(res, group)
