"""CPU oracle for the XFeat inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* of the reference algorithm (verlab/accelerated_features,
`modules/xfeat.py`, `modules/model.py`, `modules/interpolator.py`) as stateless functions over a
plain ``dict[str, Tensor]`` of the published weights.  It runs on the CPU with ATen ops -- the
same third-party arithmetic (PyTorch CPU: oneDNN / MKL) the reference itself dispatches to; the
reference pins only "pytorch >= 1.10" (README.md:90), this image has torch 2.11.0.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``accelerated_features_b200``) never does.

Parity pinning: the reference holds no tests / golden vectors for this path, so
the oracle is pinned against outputs of the *live, unmodified reference* imported from
``/root/reference`` in the build container: ``tools/make_golden.py`` generated ``tests/golden/*.npz``
and ``tests/test_oracle_golden.py`` checks this restatement against them (bit-exact for integer
outputs; float outputs agree to 0 ulp on the build container and are tested at 1e-6).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

BN_EPS = 1e-5       # nn.BatchNorm2d / BatchNorm1d default, model.py:20,99-110
IN_EPS = 1e-5       # nn.InstanceNorm2d default, model.py:35

_DEFAULT_WEIGHTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                "accelerated_features_b200", "weights", "xfeat_state.npz")


def load_state(path: str = _DEFAULT_WEIGHTS) -> State:
    """weights/xfeat.pt re-saved as npz (same keys minus num_batches_tracked); xfeat.py:30-35."""
    with np.load(path) as z:
        return {k: torch.from_numpy(z[k].copy()) for k in z.files}


# --------------------------------------------------------------------------------------------
# model.py
# --------------------------------------------------------------------------------------------
def _basic_layer(sd: State, prefix: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    """BasicLayer = Conv2d(bias=False) -> BatchNorm2d(affine=False, eval) -> ReLU; model.py:12-25."""
    x = F.conv2d(x, sd[prefix + ".layer.0.weight"], None, stride=stride, padding=padding)
    x = F.batch_norm(x, sd[prefix + ".layer.1.running_mean"], sd[prefix + ".layer.1.running_var"],
                     None, None, False, 0.0, BN_EPS)
    return F.relu(x)


def unfold8(x: torch.Tensor, ws: int = 8) -> torch.Tensor:
    """_unfold2d: out[b, ws*i+j, h, w] = x[b, 0, ws*h+i, ws*w+j]; model.py:113-120."""
    B, C, H, W = x.shape
    x = x.reshape(B, C, H // ws, ws, W // ws, ws).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(B, C * ws * ws, H // ws, W // ws)


def backbone(sd: State, x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """XFeatModel.forward with every intermediate kept; model.py:123-154."""
    out: Dict[str, torch.Tensor] = {}
    x = x.mean(dim=1, keepdim=True)                                   # model.py:135
    x = F.instance_norm(x, eps=IN_EPS)                                # model.py:136 (per-image stats)
    out["xn"] = x
    # block1, model.py:43-48,139
    t = _basic_layer(sd, "block1.0", x, 1)
    out["b1_0"] = t
    t = _basic_layer(sd, "block1.1", t, 2)
    out["b1_1"] = t
    t = _basic_layer(sd, "block1.2", t, 1)
    out["b1_2"] = t
    x1 = _basic_layer(sd, "block1.3", t, 2)
    out["x1"] = x1
    # skip1, model.py:40-41,140
    sk = F.conv2d(F.avg_pool2d(x, 4, stride=4), sd["skip1.1.weight"], sd["skip1.1.bias"])
    out["x1s"] = x1 + sk
    t = _basic_layer(sd, "block2.0", x1 + sk, 1)
    out["b2_0"] = t
    x2 = _basic_layer(sd, "block2.1", t, 1)
    out["x2"] = x2
    t = _basic_layer(sd, "block3.0", x2, 2)                            # model.py:55-59,141
    out["b3_0"] = t
    t = _basic_layer(sd, "block3.1", t, 1)
    out["b3_1"] = t
    x3 = _basic_layer(sd, "block3.2", t, 1, 0)
    out["x3"] = x3
    t = _basic_layer(sd, "block4.0", x3, 2)                            # model.py:60-64,142
    t = _basic_layer(sd, "block4.1", t, 1)
    x4 = _basic_layer(sd, "block4.2", t, 1)
    out["x4"] = x4
    t = _basic_layer(sd, "block5.0", x4, 2)                            # model.py:66-71,143
    t = _basic_layer(sd, "block5.1", t, 1)
    t = _basic_layer(sd, "block5.2", t, 1)
    x5 = _basic_layer(sd, "block5.3", t, 1, 0)
    out["x5"] = x5
    # pyramid fusion, model.py:146-148
    size = (x3.shape[-2], x3.shape[-1])
    x4u = F.interpolate(x4, size, mode="bilinear")
    x5u = F.interpolate(x5, size, mode="bilinear")
    fin = x3 + x4u + x5u
    out["fusion_in"] = fin
    t = _basic_layer(sd, "block_fusion.0", fin, 1)
    t = _basic_layer(sd, "block_fusion.1", t, 1)
    feats = F.conv2d(t, sd["block_fusion.2.weight"], sd["block_fusion.2.bias"])
    out["feats"] = feats
    # heads, model.py:79-92,151-152
    t = _basic_layer(sd, "heatmap_head.0", feats, 1, 0)
    t = _basic_layer(sd, "heatmap_head.1", t, 1, 0)
    out["reliability"] = torch.sigmoid(F.conv2d(t, sd["heatmap_head.2.weight"], sd["heatmap_head.2.bias"]))
    t = unfold8(x, 8)
    out["unfold"] = t
    t = _basic_layer(sd, "keypoint_head.0", t, 1, 0)
    t = _basic_layer(sd, "keypoint_head.1", t, 1, 0)
    t = _basic_layer(sd, "keypoint_head.2", t, 1, 0)
    out["kpt_logits"] = F.conv2d(t, sd["keypoint_head.3.weight"], sd["keypoint_head.3.bias"])
    return out


def fine_matcher(sd: State, x: torch.Tensor) -> torch.Tensor:
    """fine_matcher MLP 128->512->512->512->512->64, BatchNorm1d(affine=False) eval; model.py:97-111."""
    for i, bn in ((0, 1), (3, 4), (6, 7), (9, 10)):
        x = F.linear(x, sd[f"fine_matcher.{i}.weight"], sd[f"fine_matcher.{i}.bias"])
        x = F.batch_norm(x, sd[f"fine_matcher.{bn}.running_mean"], sd[f"fine_matcher.{bn}.running_var"],
                         None, None, False, 0.0, BN_EPS)
        x = F.relu(x)
    return F.linear(x, sd["fine_matcher.12.weight"], sd["fine_matcher.12.bias"])


# --------------------------------------------------------------------------------------------
# interpolator.py
# --------------------------------------------------------------------------------------------
def sample_sparse(x: torch.Tensor, pos: torch.Tensor, H: int, W: int, mode: str) -> torch.Tensor:
    """InterpolateSparse2d.forward: grid = 2*pos/(W-1,H-1)-1, grid_sample(align_corners=False);
    interpolator.py:17-33.  Returns (B,N,C)."""
    grid = 2.0 * (pos / torch.tensor([W - 1, H - 1], dtype=pos.dtype)) - 1.0
    grid = grid.unsqueeze(-2).to(x.dtype)
    x = F.grid_sample(x, grid, mode=mode, align_corners=False)
    return x.permute(0, 2, 3, 1).squeeze(-2)


def source_coord(p: np.ndarray, size_pos: int, size_map: int) -> np.ndarray:
    """Explicit fp32 restatement of the coordinate the sampler reads for integer position p:
    normgrid (interpolator.py:17-19) then ATen grid_sampler_unnormalize(align_corners=False):
    ((g + 1) * size_map - 1) / 2.  Used by tests to pin the convention (~ p*size_map/(size_pos-1) - 0.5)."""
    p = p.astype(np.float32)
    g = np.float32(2.0) * (p / np.float32(size_pos - 1)) - np.float32(1.0)
    return ((g + np.float32(1.0)) * np.float32(size_map) - np.float32(1.0)) / np.float32(2.0)


# --------------------------------------------------------------------------------------------
# xfeat.py -- sparse path
# --------------------------------------------------------------------------------------------
def parse_input(x):
    """xfeat.py:396-403: 3-D -> add batch dim; numpy (B,H,W,C) -> tensor (B,C,H,W)/255."""
    if len(x.shape) == 3:
        x = x[None, ...]
    if isinstance(x, np.ndarray):
        x = torch.tensor(x).permute(0, 3, 1, 2) / 255
    return x


def preprocess_tensor(x) -> Tuple[torch.Tensor, float, float]:
    """xfeat.py:219-240: to float, resize (bilinear, align_corners=False) to multiples of 32."""
    if isinstance(x, np.ndarray):
        if x.ndim == 3:
            x = torch.tensor(x).permute(2, 0, 1)[None]
        elif x.ndim == 2:
            x = torch.tensor(x[..., None]).permute(2, 0, 1)[None]
        else:
            raise RuntimeError("For numpy arrays, only (H,W) or (H,W,C) format is supported.")
    if len(x.shape) != 4:
        raise RuntimeError("Input tensor needs to be in (B,C,H,W) format")
    x = x.float()
    H, W = x.shape[-2:]
    _H, _W = (H // 32) * 32, (W // 32) * 32
    rh, rw = H / _H, W / _W
    x = F.interpolate(x, (_H, _W), mode="bilinear", align_corners=False)
    return x, rh, rw


def kpts_heatmap(kpt_logits: torch.Tensor, softmax_temp: float = 1.0) -> torch.Tensor:
    """xfeat.py:242-247: softmax over 65, drop dustbin, heat[b,0,8h+i,8w+j] = p[b,8i+j,h,w]."""
    p = F.softmax(kpt_logits * softmax_temp, 1)[:, :64]
    B, _, H, W = p.shape
    p = p.permute(0, 2, 3, 1).reshape(B, H, W, 8, 8)
    return p.permute(0, 1, 3, 2, 4).reshape(B, 1, H * 8, W * 8)


def nms(heat: torch.Tensor, threshold: float = 0.05, kernel_size: int = 5) -> torch.Tensor:
    """xfeat.py:249-263: 5x5 max-pool equality + threshold; raster-order (x,y) int64, zero padded."""
    B = heat.shape[0]
    local_max = F.max_pool2d(heat, kernel_size, stride=1, padding=kernel_size // 2)
    keep = (heat == local_max) & (heat > threshold)
    per_img = [k.nonzero()[..., 1:].flip(-1) for k in keep]
    n = max(len(p) for p in per_img)
    pos = torch.zeros((B, n, 2), dtype=torch.long)
    for b, p in enumerate(per_img):
        pos[b, : len(p)] = p
    return pos


def nms_counts(heat: torch.Tensor, threshold: float = 0.05, kernel_size: int = 5) -> List[int]:
    local_max = F.max_pool2d(heat, kernel_size, stride=1, padding=kernel_size // 2)
    keep = (heat == local_max) & (heat > threshold)
    return [int(k.sum()) for k in keep]


def sparse_scores(heat: torch.Tensor, reliability: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """xfeat.py:77-80: nearest(K1h) * bilinear(H1); padding rows (0,0) -> -1."""
    _, _, H, W = heat.shape
    s = (sample_sparse(heat, pos, H, W, "nearest") * sample_sparse(reliability, pos, H, W, "bilinear")).squeeze(-1)
    s[torch.all(pos == 0, dim=-1)] = -1
    return s


def select_topk(pos: torch.Tensor, scores: torch.Tensor, top_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """xfeat.py:83-87: argsort(-scores) (unstable), gather, keep top_k."""
    idxs = torch.argsort(-scores)
    px = torch.gather(pos[..., 0], -1, idxs)[:, :top_k]
    py = torch.gather(pos[..., 1], -1, idxs)[:, :top_k]
    return torch.cat([px[..., None], py[..., None]], dim=-1), torch.gather(scores, -1, idxs)[:, :top_k]


def detect_and_compute(sd: State, x, top_k: int = 4096, detection_threshold: float = 0.05,
                       return_stages: bool = False):
    """XFeat.detectAndCompute; xfeat.py:49-103."""
    x, rh, rw = preprocess_tensor(x)
    B, _, H, W = x.shape
    st = backbone(sd, x)
    M = F.normalize(st["feats"], dim=1)                                  # xfeat.py:70
    heat = kpts_heatmap(st["kpt_logits"])                                # xfeat.py:73
    pos = nms(heat, detection_threshold, 5)                              # xfeat.py:74
    scores_all = sparse_scores(heat, st["reliability"], pos)             # xfeat.py:77-80
    kp, scores = select_topk(pos, scores_all, top_k)                     # xfeat.py:83-87
    desc = sample_sparse(M, kp, H, W, "bicubic")                         # xfeat.py:90
    desc = F.normalize(desc, dim=-1)                                     # xfeat.py:93
    kpf = kp * torch.tensor([rw, rh]).view(1, 1, -1)                     # xfeat.py:96
    valid = scores > 0                                                   # xfeat.py:98
    res = [{"keypoints": kpf[b][valid[b]], "scores": scores[b][valid[b]], "descriptors": desc[b][valid[b]]}
           for b in range(B)]
    if return_stages:
        st.update(x=x, M=M, heat=heat, nms_pos=pos, scores_all=scores_all, kp_int=kp, scores_topk=scores,
                  desc_topk=desc, rh=rh, rw=rw)
        return res, st
    return res


def mnn_match(f1: torch.Tensor, f2: torch.Tensor, min_cossim: float = 0.82) -> Tuple[torch.Tensor, torch.Tensor]:
    """XFeat.match; xfeat.py:327-348.  argmax ties -> first index (torch.max CPU)."""
    s = f1 @ f2.t()
    st = f2 @ f1.t()
    _, m12 = s.max(dim=1)
    _, m21 = st.max(dim=1)
    idx0 = torch.arange(len(m12))
    mutual = m21[m12] == idx0
    if min_cossim > 0:
        best, _ = s.max(dim=1)
        good = best > min_cossim
        return idx0[mutual & good], m12[mutual & good]
    return idx0[mutual], m12[mutual]


def mnn_ambiguity(f1: torch.Tensor, f2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Test helper (not in the reference): top-1/top-2 gap per row and per column of f1 @ f2^T.
    Rows/cols whose gap is below fp32 accumulation noise may legitimately resolve differently
    under a different summation order (SURVEY.md section 7 hard part 1)."""
    s = (f1.double() @ f2.double().t())
    r = torch.topk(s, 2, dim=1).values
    c = torch.topk(s, 2, dim=0).values
    return (r[:, 0] - r[:, 1]).float(), (c[0] - c[1]).float()


def match_xfeat(sd: State, img1, img2, top_k: int = 4096, min_cossim: float = -1):
    """XFeat.match_xfeat; xfeat.py:165-186 (B=1 semantics: only batch item 0 is used)."""
    o1 = detect_and_compute(sd, parse_input(img1), top_k)[0]
    o2 = detect_and_compute(sd, parse_input(img2), top_k)[0]
    i0, i1 = mnn_match(o1["descriptors"], o2["descriptors"], min_cossim)
    return o1["keypoints"][i0].numpy(), o2["keypoints"][i1].numpy()


# --------------------------------------------------------------------------------------------
# xfeat.py -- semi-dense (XFeat*) path
# --------------------------------------------------------------------------------------------
def create_xy(h: int, w: int) -> torch.Tensor:
    """xfeat.py:350-354: (h*w, 2) int64 (x,y), y-major."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return torch.cat([x[..., None], y[..., None]], -1).reshape(-1, 2)


def extract_dense(sd: State, x: torch.Tensor, top_k: int = 8000, return_stages: bool = False):
    """XFeat.extractDense; xfeat.py:356-377.  Descriptors are NOT normalised."""
    if top_k < 1:
        top_k = 100_000_000
    x, rh, rw = preprocess_tensor(x)
    st = backbone(sd, x)
    M, Hm = st["feats"], st["reliability"]
    B, C, h, w = M.shape
    xy = (create_xy(h, w) * 8).expand(B, -1, -1)
    Mf = M.permute(0, 2, 3, 1).reshape(B, -1, C)
    Hf = Hm.permute(0, 2, 3, 1).reshape(B, -1)
    vals, idx = torch.topk(Hf, k=min(len(Hf[0]), top_k), dim=-1)
    feats = torch.gather(Mf, 1, idx[..., None].expand(-1, -1, 64))
    kp = torch.gather(xy, 1, idx[..., None].expand(-1, -1, 2))
    kp = kp * torch.tensor([rw, rh]).view(1, -1)
    if return_stages:
        return kp, feats, dict(st, x=x, topk_idx=idx, topk_val=vals, rel_flat=Hf)
    return kp, feats


def extract_dualscale(sd: State, x: torch.Tensor, top_k: int, s1: float = 0.6, s2: float = 1.3):
    """XFeat.extract_dualscale; xfeat.py:379-394."""
    x1 = F.interpolate(x, scale_factor=s1, align_corners=False, mode="bilinear")
    x2 = F.interpolate(x, scale_factor=s2, align_corners=False, mode="bilinear")
    kp1, f1 = extract_dense(sd, x1, int(top_k * 0.20))
    kp2, f2 = extract_dense(sd, x2, int(top_k * 0.80))
    kp = torch.cat([kp1 / s1, kp2 / s2], dim=1)
    sc = torch.cat([torch.ones(kp1.shape[:2]) * (1 / s1), torch.ones(kp2.shape[:2]) * (1 / s2)], dim=1)
    return kp, sc, torch.cat([f1, f2], dim=1)


def detect_and_compute_dense(sd: State, x, top_k: int = 4096, multiscale: bool = True):
    """XFeat.detectAndComputeDense; xfeat.py:105-128."""
    if multiscale:
        kp, sc, feats = extract_dualscale(sd, x, top_k)
    else:
        kp, feats = extract_dense(sd, x, top_k)
        sc = torch.ones(kp.shape[:2])
    return {"keypoints": kp, "descriptors": feats, "scales": sc}


def batch_match(f1: torch.Tensor, f2: torch.Tensor, min_cossim: float = -1):
    """XFeat.batch_match; xfeat.py:265-290: one bmm, argmax both ways on the same matrix."""
    s = torch.bmm(f1, f2.permute(0, 2, 1))
    m12 = torch.argmax(s, dim=-1)
    m21 = torch.argmax(s.permute(0, 2, 1), dim=-1)
    idx0 = torch.arange(len(m12[0]))
    out = []
    for b in range(len(f1)):
        mutual = m21[b][m12[b]] == idx0
        if min_cossim > 0:
            best, _ = s[b].max(dim=1)
            good = best > min_cossim
            out.append((idx0[mutual & good], m12[b][mutual & good]))
        else:
            out.append((idx0[mutual], m12[b][mutual]))
    return out


def subpix_softmax2d(heatmaps: torch.Tensor, temp: float = 3) -> torch.Tensor:
    """xfeat.py:292-304: expectation of (x - W//2, y - H//2) under softmax(temp * logits), x fastest."""
    N, H, W = heatmaps.shape
    p = torch.softmax(temp * heatmaps.view(-1, H * W), -1).view(-1, H, W)
    x, y = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    x = x - (W // 2)
    y = y - (H // 2)
    c = torch.cat([(x[None] * p)[..., None], (y[None] * p)[..., None]], -1).view(N, H * W, 2)
    return c.sum(1)


def refine_matches(sd: State, d0, d1, matches, batch_idx: int, fine_conf: float = 0.25,
                   return_stages: bool = False):
    """XFeat.refine_matches; xfeat.py:306-325."""
    idx0, idx1 = matches[batch_idx]
    f1 = d0["descriptors"][batch_idx][idx0]
    f2 = d1["descriptors"][batch_idx][idx1]
    k0 = d0["keypoints"][batch_idx][idx0]
    k1 = d1["keypoints"][batch_idx][idx1]
    sc0 = d0["scales"][batch_idx][idx0]
    logits = fine_matcher(sd, torch.cat([f1, f2], dim=-1))
    conf = F.softmax(logits * 3, dim=-1).max(dim=-1)[0]
    off = subpix_softmax2d(logits.view(-1, 8, 8))
    k0 = k0 + off * sc0[:, None]
    good = conf > fine_conf
    res = torch.cat([k0[good], k1[good]], dim=-1)
    if return_stages:
        return res, dict(logits=logits, conf=conf, offsets=off, good=good)
    return res


def match_xfeat_star(sd: State, im_set1, im_set2, top_k: int = 4096):
    """XFeat.match_xfeat_star; xfeat.py:188-217.  Return type switches on B exactly as the reference."""
    im_set1 = parse_input(im_set1)
    im_set2 = parse_input(im_set2)
    o1 = detect_and_compute_dense(sd, im_set1, top_k)
    o2 = detect_and_compute_dense(sd, im_set2, top_k)
    idxs = batch_match(o1["descriptors"], o2["descriptors"])
    B = len(im_set1)
    matches = [refine_matches(sd, o1, o2, idxs, b) for b in range(B)]
    return matches if B > 1 else (matches[0][:, :2].numpy(), matches[0][:, 2:].numpy())


# --------------------------------------------------------------------------------------------
# Explicit (numpy) restatements of the integer / index stages.  These state the rules the CUDA
# kernels implement without leaning on ATen, and are cross-checked against the ATen versions above.
# --------------------------------------------------------------------------------------------
def nms_numpy(heat: np.ndarray, threshold: float = 0.05, r: int = 2) -> np.ndarray:
    """(H,W) float32 -> (N,2) int64 (x,y) raster order.  -inf padding (MaxPool2d), x == max, x > thr."""
    H, W = heat.shape
    pad = np.full((H + 2 * r, W + 2 * r), -np.inf, np.float32)
    pad[r:r + H, r:r + W] = heat
    m = np.full((H, W), -np.inf, np.float32)
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            m = np.maximum(m, pad[dy:dy + H, dx:dx + W])
    ys, xs = np.nonzero((heat == m) & (heat > np.float32(threshold)))
    return np.stack([xs, ys], -1).astype(np.int64)


def mnn_numpy(s: np.ndarray, min_cossim: float = -1.0) -> Tuple[np.ndarray, np.ndarray]:
    """Mutual-NN on a materialised similarity matrix; first-index ties (np.argmax)."""
    m12 = s.argmax(1)
    m21 = s.argmax(0)
    i = np.arange(s.shape[0])
    keep = m21[m12] == i
    if min_cossim > 0:
        keep &= s.max(1) > np.float32(min_cossim)
    return i[keep].astype(np.int64), m12[keep].astype(np.int64)


def canonical_topk_order(scores: np.ndarray, lin_idx: np.ndarray) -> np.ndarray:
    """Deterministic order the CUDA top-k uses: score descending, then linear pixel index ascending.
    The reference's argsort is unstable, so tests compare after this canonicalisation."""
    return np.lexsort((lin_idx, -scores.astype(np.float64)))
