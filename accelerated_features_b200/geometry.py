"""Geometric verification of matches on the GPU: the step every caller of the reference runs right after
matching -- `cv2.findHomography(points1, points2, cv2.USAC_MAGSAC, thr, maxIters=700, confidence=0.995)` in
realtime_demo.py:225 and the notebooks -- for a whole batch of pairs in two kernel launches (csrc/ransac.cu), without the
matches leaving the device."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


def find_homography_batch(pts0: torch.Tensor, pts1: torch.Tensor, counts: Optional[torch.Tensor] = None, thr: float = 3.0,
                          iters: int = 1024, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """pts0, pts1: (B, n_max, 2) float32 CUDA tensors of matched coordinates (pair b uses the first counts[b] rows; None: all).
    Returns H (B,3,3) float32 with H[2,2] = 1 mapping image-0 to image-1 pixels, inliers (B, n_max) bool, n_inliers (B) int32.
    `thr`: forward transfer error in pixels of image 1 (cv2's ransacReprojThreshold)."""
    lib = _lib.load()
    if not (pts0.is_cuda and pts1.is_cuda):
        raise RuntimeError("find_homography_batch runs on the GPU only (no CPU fallback)")
    p0 = pts0.float().contiguous()
    p1 = pts1.float().contiguous()
    if p0.ndim != 3 or p0.shape != p1.shape or p0.shape[-1] != 2:
        raise RuntimeError("points must be two (B, n, 2) tensors of the same shape")
    B, n_max, _ = p0.shape
    dev = p0.device
    H = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((B, n_max), dtype=torch.uint8, device=dev)
    n_inl = torch.empty((B,), dtype=torch.int32, device=dev)
    if B == 0 or n_max == 0:
        return H, mask.bool(), n_inl
    cnt = None if counts is None else counts.to(dev).to(torch.int32).contiguous()
    ws = torch.empty(max(256, lib.xfeat_ransac_workspace_bytes(B, iters)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xfeat_ransac_homography(p0.data_ptr(), p1.data_ptr(), None if cnt is None else cnt.data_ptr(), n_max, B,
                                               float(thr), int(iters), int(seed) & 0xffffffff, H.data_ptr(), mask.data_ptr(),
                                               n_inl.data_ptr(), ws.data_ptr(), ws.numel(),
                                               torch.cuda.current_stream(dev).cuda_stream), "xfeat_ransac_homography")
    return H, mask.bool(), n_inl


def find_homography(points1, points2, thr: float = 3.0, iters: int = 1024, seed: int = 0, device: Optional[torch.device] = None):
    """cv2.findHomography-shaped call for ONE pair: (N,2) arrays -> (H (3,3) float64 or None, mask (N,1) uint8)."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    p0 = torch.as_tensor(np.asarray(points1, dtype=np.float32)).reshape(1, -1, 2).to(dev)
    p1 = torch.as_tensor(np.asarray(points2, dtype=np.float32)).reshape(1, -1, 2).to(dev)
    n = p0.shape[1]
    if n < 4:
        return None, np.zeros((n, 1), np.uint8)
    H, mask, n_inl = find_homography_batch(p0, p1, None, thr, iters, seed)
    if int(n_inl.item()) < 4:
        return None, np.zeros((n, 1), np.uint8)
    return H[0].double().cpu().numpy(), mask[0].to(torch.uint8).cpu().numpy().reshape(-1, 1)


def find_essential_batch(x0: torch.Tensor, x1: torch.Tensor, counts: Optional[torch.Tensor] = None, thr: float = 1e-3,
                         iters: int = 1024, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Essential-matrix RANSAC for a batch of pairs (the relative-pose step of the 1500-pair benchmarks,
    modules/eval/megadepth1500.py:98-113).  x0, x1: (B, n_max, 2) NORMALISED image coordinates (K^-1 applied) on the GPU;
    `thr`: Sampson distance threshold in the same units (pixels / focal length).  Returns E (B,3,3) with x1^T E x0 = 0,
    inliers (B, n_max) bool, n_inliers (B)."""
    lib = _lib.load()
    if not (x0.is_cuda and x1.is_cuda):
        raise RuntimeError("find_essential_batch runs on the GPU only (no CPU fallback)")
    a, b = x0.float().contiguous(), x1.float().contiguous()
    if a.ndim != 3 or a.shape != b.shape or a.shape[-1] != 2:
        raise RuntimeError("points must be two (B, n, 2) tensors of the same shape")
    B, n_max, _ = a.shape
    dev = a.device
    E = torch.zeros((B, 3, 3), dtype=torch.float32, device=dev)
    mask = torch.zeros((B, n_max), dtype=torch.uint8, device=dev)
    n_inl = torch.zeros((B,), dtype=torch.int32, device=dev)
    if B == 0 or n_max == 0:
        return E, mask.bool(), n_inl
    cnt = None if counts is None else counts.to(dev).to(torch.int32).contiguous()
    ws = torch.empty(max(256, lib.xfeat_ransac_workspace_bytes(B, iters)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xfeat_ransac_essential(a.data_ptr(), b.data_ptr(), None if cnt is None else cnt.data_ptr(), n_max, B,
                                              float(thr), int(iters), int(seed) & 0xffffffff, E.data_ptr(), mask.data_ptr(),
                                              n_inl.data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream(dev).cuda_stream), "xfeat_ransac_essential")
    return E, mask.bool(), n_inl


def recover_pose(E: np.ndarray, x0: np.ndarray, x1: np.ndarray):
    """(R, t) with X1 = R X0 + t from an essential matrix and normalised inlier correspondences: the candidate (of the four
    decompositions) that puts the most triangulated points in front of both cameras (what cv2.recoverPose does)."""
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    Wm = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    best, ret = -1, None
    h0 = np.concatenate([x0, np.ones((len(x0), 1))], 1)
    h1 = np.concatenate([x1, np.ones((len(x1), 1))], 1)
    for R in (U @ Wm @ Vt, U @ Wm.T @ Vt):
        for t in (U[:, 2], -U[:, 2]):
            # depth of X0 along ray h0: solve z0 (R h0) - z1 h1 = -t in the least-squares sense, per point
            a = (R @ h0.T).T
            A = np.stack([a, -h1], 2)                                    # (n, 3, 2)
            AtA = np.einsum("nij,nik->njk", A, A)
            Atb = np.einsum("nij,i->nj", A, -t)
            z = np.linalg.solve(AtA + 1e-12 * np.eye(2), Atb[..., None])[..., 0]
            good = int(((z[:, 0] > 0) & (z[:, 1] > 0)).sum())
            if good > best:
                best, ret = good, (R, t)
    return ret
