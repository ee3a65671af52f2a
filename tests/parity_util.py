"""Helpers of the end-to-end parity tests: EXACT set comparison against the oracle, with the only admitted differences being
an enumerated set of near-ties: a keypoint / match may differ between the CUDA path and the oracle only where the
oracle's own decision hangs on a float comparison whose margin is below the accumulated fp32 re-association error of the
backbone (heat-map 5e-5, scores 1e-6, descriptors 2e-5 -- the stage tests assert those bounds)."""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, List, Set, Tuple

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_RECORDS: Dict[str, dict] = {}


def record(name: str, **kv):
    """Observed parity counts; dumped to gpurun_out/parity.json (copied to profiles/rNN/parity.json by hand)."""
    _RECORDS[name] = {k: (v if isinstance(v, (int, float, str, list, bool)) else str(v)) for k, v in kv.items()}
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity.json")
        old = {}
        if os.path.exists(path):
            with open(path) as f:
                old = json.load(f)
        old.update(_RECORDS)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except Exception:
        pass
    print(f"[parity] {name}: {_RECORDS[name]}")


def kp_set(kp) -> Set[Tuple[float, float]]:
    return {(float(x), float(y)) for x, y in np.asarray(kp)}


def keypoint_margins(st: dict, b: int, pts: Iterable[Tuple[int, int]], top_k: int, thr: float = 0.05) -> List[float]:
    """For integer pixel positions `pts` of image b: the smallest margin of any float comparison that decides whether the
    ORACLE keeps the point (threshold test, 5x5 NMS equality, score > 0, top-k cut).  st = stages of orc.detect_and_compute."""
    heat = st["heat"][b, 0].numpy()
    H, W = heat.shape
    scores_topk = st["scores_topk"][b].numpy()
    kth = float(scores_topk[min(top_k, len(scores_topk)) - 1]) if len(scores_topk) else 0.0
    saturated = (st["scores_all"][b] > 0).sum().item() >= top_k
    rel = st["reliability"][b:b + 1]
    out = []
    for (x, y) in pts:
        v = float(heat[y, x])
        y0, y1, x0, x1 = max(0, y - 2), min(H, y + 3), max(0, x - 2), min(W, x + 3)
        win = heat[y0:y1, x0:x1].copy()
        win[y - y0, x - x0] = -np.inf
        m = [abs(v - thr), abs(v - float(win.max()))]
        # score of this pixel under the oracle's own formula
        from oracle import xfeat_oracle as orc
        pos = torch.tensor([[[x, y]]], dtype=torch.long)
        s = float(orc.sparse_scores(st["heat"][b:b + 1], rel, pos)[0, 0]) if not (x == 0 and y == 0) else -1.0
        m.append(abs(s))                       # `scores > 0`
        if saturated:
            m.append(abs(s - kth))             # top-k cut
        out.append(min(m))
    return out


def assert_keypoints_exact_modulo_ties(name: str, got_kp: np.ndarray, want_kp: np.ndarray, st: dict, b: int, top_k: int,
                                       rw: float = 1.0, rh: float = 1.0, eps: float = 2e-4, max_diff: int = 8) -> int:
    """Exact set equality of keypoints; every differing keypoint must have an oracle decision margin < eps."""
    g, w = kp_set(got_kp), kp_set(want_kp)
    diff = g ^ w
    if diff:
        assert len(diff) <= max_diff, f"{name}: {len(diff)} differing keypoints (limit {max_diff})"
        pts = [(int(round(x / rw)), int(round(y / rh))) for x, y in diff]
        margins = keypoint_margins(st, b, pts, top_k)
        for p, mg in zip(pts, margins):
            assert mg < eps, f"{name}: keypoint {p} differs although its decision margin is {mg:.3e} (eps {eps:.1e})"
    return len(diff)


def match_set(mk0, mk1) -> Set[Tuple[float, float, float, float]]:
    return {(float(a), float(b), float(c), float(d)) for (a, b), (c, d) in zip(np.asarray(mk0), np.asarray(mk1))}


def assert_matches_exact_modulo_ties(name: str, got: Set[tuple], want: Set[tuple], want_kp0: np.ndarray, want_desc0: torch.Tensor,
                                     want_kp1: np.ndarray, want_desc1: torch.Tensor, kp_diff0: int = 0, kp_diff1: int = 0,
                                     eps: float = 1e-4, max_diff: int = 8) -> int:
    """Exact equality of the matched coordinate pairs.  A differing pair must touch a row / column of the oracle's similarity
    matrix whose top-1 / top-2 gap is below eps (descriptors agree to 2e-5, so dot products to ~4e-5), or a keypoint that
    itself differs (kp_diff > 0: then each differing keypoint can change up to two pairs)."""
    diff = got ^ want
    if not diff:
        return 0
    assert len(diff) <= max_diff + 2 * (kp_diff0 + kp_diff1), f"{name}: {len(diff)} differing matches"
    i0 = {(float(x), float(y)): i for i, (x, y) in enumerate(want_kp0)}
    i1 = {(float(x), float(y)): i for i, (x, y) in enumerate(want_kp1)}
    s = want_desc0.double() @ want_desc1.double().t()
    for (a, b, c, d) in diff:
        ia, ib = i0.get((a, b)), i1.get((c, d))
        if ia is None or ib is None:
            assert kp_diff0 + kp_diff1 > 0, f"{name}: match {(a, b, c, d)} uses a keypoint the oracle does not have"
            continue
        r = torch.topk(s[ia], 2).values
        cc = torch.topk(s[:, ib], 2).values
        gap = min(float(r[0] - r[1]), float(cc[0] - cc[1]))
        assert gap < eps or kp_diff0 + kp_diff1 > 0, f"{name}: robust match {(a, b, c, d)} differs (gap {gap:.3e})"
    return len(diff)
