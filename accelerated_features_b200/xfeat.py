"""Drop-in replacement for `modules.xfeat.XFeat` (verlab/accelerated_features) on NVIDIA B200.

Same constructor, method names, argument meaning, return types and error behaviour as the reference class
(modules/xfeat.py), but every operation of the hot path runs in libxfeat_sm100.so (hand-written sm_100a CUDA kernels
behind the C-ABI of include/xfeat_b200.h).  PyTorch is used for device memory, streams and the final slicing only.
There is no CPU / eager fallback: without a CUDA device and the built library the constructor raises.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from . import weights as _weights

_F32 = np.float32


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Net:
    """Stands in for `XFeat.net` (XFeatModel, model.py:27): callable (B,C,H,W) -> (feats, keypoints, heatmap)."""

    def __init__(self, owner: "XFeat"):
        self._o = owner

    @torch.inference_mode()
    def __call__(self, x: torch.Tensor):
        """XFeatModel.forward (model.py:123-154).  x: (B,C,H,W) with H, W multiples of 32.
        Returns feats (B,64,H/8,W/8), keypoint logits (B,65,H/8,W/8), reliability (B,1,H/8,W/8) as NCHW *views*
        of channels-last storage."""
        o = self._o
        x = x.to(o.dev)
        B, C, H, W = x.shape
        if H % 32 or W % 32:
            raise RuntimeError("XFeatModel input must have H, W divisible by 32")
        xn = o._preprocess(x, H, W, div255=False)
        feats, heat, rel, logits = o._run_net(xn, B, H, W, want_logits=True)
        return feats.permute(0, 3, 1, 2), logits.permute(0, 3, 1, 2), rel.unsqueeze(1)

    def fine_matcher(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("fine_matcher runs fused inside XFeat.match_xfeat_star (xfeat_refine)")


class XFeat:
    """B200-native XFeat inference (sparse `detectAndCompute` / `match_xfeat`, semi-dense `match_xfeat_star`)."""

    def __init__(self, weights=_weights.DEFAULT_WEIGHTS, top_k: int = 4096, detection_threshold: float = 0.05,
                 device: Optional[int] = None):
        # reference: xfeat.py:23-46.  `weights`: path (.pt/.npz), state_dict mapping, or None (random init)
        if not torch.cuda.is_available():
            raise RuntimeError("accelerated_features_b200.XFeat needs a CUDA device (sm_100a); there is no CPU fallback")
        self._lib = _lib.load()
        dev_index = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", dev_index)
        self.top_k = top_k
        self.detection_threshold = detection_threshold
        sd = _weights.random_state_dict(0) if weights is None else _weights.load_state_dict(weights)
        if isinstance(weights, str):
            print("loading weights from: " + weights)
        blob = _weights.pack_weights(sd)
        import ctypes
        handle = ctypes.c_void_p()
        _lib.check(self._lib.xfeat_create(ctypes.byref(handle), dev_index, blob.ctypes.data, blob.size), "xfeat_create")
        self._ctx = handle
        self._ws: Optional[torch.Tensor] = None
        self.net = _Net(self)
        self.interpolator = "bicubic"   # reference keeps an InterpolateSparse2d('bicubic') here (xfeat.py:37)
        self.kornia_available = False
        self.lighterglue = None

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                self._lib.xfeat_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    # ------------------------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=self.dev)
        return self._ws

    def _empty(self, shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.dev)

    @staticmethod
    def _img_args(x: torch.Tensor):
        """(tensor, dtype code) with a layout the kernels can address in place."""
        if x.dtype == torch.uint8:
            return x, 1
        if x.dtype != torch.float32:
            x = x.float()          # same as the reference's `.float()` (xfeat.py:233)
        return x, 0

    def _preprocess(self, x: torch.Tensor, H: int, W: int, div255: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """xfeat_preprocess: resize to (H,W), channel mean, InstanceNorm -> (B,H,W) fp32 (optionally into `out`)."""
        x, code = self._img_args(x)
        B, C, Hi, Wi = x.shape
        sb, sc, sh, sw = x.stride()
        xn = self._empty((B, H, W)) if out is None else out
        stats = self._empty((B, 2), torch.float64)
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_preprocess(x.data_ptr(), code, B, C, Hi, Wi, sb, sc, sh, sw, int(div255), H, W,
                                                  xn.data_ptr(), stats.data_ptr(), self._stream()), "xfeat_preprocess")
        return xn

    def _run_net(self, xn: torch.Tensor, B: int, H: int, W: int, want_logits: bool = False):
        feats = self._empty((B, H // 8, W // 8, 64))
        heat = self._empty((B, H, W))
        rel = self._empty((B, H // 8, W // 8))
        logits = self._empty((B, H // 8, W // 8, 65)) if want_logits else None
        nbytes = self._lib.xfeat_net_workspace_bytes(B, H, W)
        ws = self._workspace(nbytes)
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_net(self._ctx, xn.data_ptr(), B, H, W, feats.data_ptr(), heat.data_ptr(),
                                           rel.data_ptr(), _ptr(logits), ws.data_ptr(), ws.numel(), self._stream()),
                       "xfeat_net")
        return feats, heat, rel, logits

    def _to_bchw(self, x) -> torch.Tensor:
        """preprocess_tensor's input handling (xfeat.py:221-233): numpy (H,W)/(H,W,C) or tensor (B,C,H,W)."""
        if isinstance(x, np.ndarray):
            if len(x.shape) == 3:
                x = torch.from_numpy(np.ascontiguousarray(x)).permute(2, 0, 1)[None]
            elif len(x.shape) == 2:
                x = torch.from_numpy(np.ascontiguousarray(x))[None, None]
            else:
                raise RuntimeError('For numpy arrays, only (H,W) or (H,W,C) format is supported.')
        if len(x.shape) != 4:
            raise RuntimeError('Input tensor needs to be in (B,C,H,W) format')
        return x.to(self.dev, non_blocking=True)

    # ------------------------------------------------------------------------------------------------------------
    # sparse path
    # ------------------------------------------------------------------------------------------------------------
    def _detect_sparse_device(self, x, top_k: int, detection_threshold: float, div255: bool = False):
        """Whole sparse extraction on the device, fixed-capacity outputs, no host sync.  `x` is one image batch or a list
        of batches of identical shape (they are normalised into one activation batch without concatenating the inputs).
        Returns dict of device tensors: keypoints (B,k,2), scores (B,k), descriptors (B,k,64), n_valid (B) int32."""
        xs = [self._to_bchw(t) for t in (x if isinstance(x, (list, tuple)) else [x])]
        _, _, Hi, Wi = xs[0].shape
        if any(t.shape[1:] != xs[0].shape[1:] for t in xs):
            raise RuntimeError("image batches must share (C,H,W)")
        B = sum(t.shape[0] for t in xs)
        H, W = (Hi // 32) * 32, (Wi // 32) * 32
        if H == 0 or W == 0:
            raise RuntimeError("image smaller than 32 pixels")
        rh, rw = Hi / H, Wi / W                                     # python floats, as xfeat.py:237
        xn = self._empty((B, H, W))
        o = 0
        for t in xs:
            self._preprocess(t, H, W, div255, out=xn[o:o + t.shape[0]])
            o += t.shape[0]
        feats, heat, rel, _ = self._run_net(xn, B, H, W)
        kpts = self._empty((B, top_k, 2))
        scores = self._empty((B, top_k))
        desc = self._empty((B, top_k, 64))
        n_valid = self._empty((B,), torch.int32)
        n_cand = self._empty((B,), torch.int32)
        nbytes = self._lib.xfeat_sparse_workspace_bytes(B, H, W, top_k)
        if nbytes == 0:
            raise _lib.XFeatLibraryError("xfeat_sparse_workspace_bytes failed: " + self._lib.xfeat_last_error().decode())
        ws = self._workspace(nbytes)
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_detect_sparse(self._ctx, feats.data_ptr(), heat.data_ptr(), rel.data_ptr(), B, H, W,
                                                     top_k, float(detection_threshold), float(_F32(rw)), float(_F32(rh)),
                                                     kpts.data_ptr(), scores.data_ptr(), desc.data_ptr(),
                                                     n_valid.data_ptr(), n_cand.data_ptr(), None, ws.data_ptr(),
                                                     ws.numel(), self._stream()), "xfeat_detect_sparse")
        return {"keypoints": kpts, "scores": scores, "descriptors": desc, "n_valid": n_valid, "n_cand": n_cand,
                "feats": feats, "heat": heat, "reliability": rel, "H": H, "W": W}

    @torch.inference_mode()
    def detectAndCompute(self, x, top_k=None, detection_threshold=None) -> List[Dict[str, torch.Tensor]]:
        """Compute sparse keypoints & descriptors, batched (reference: xfeat.py:49-103).

        x -> torch.Tensor(B,C,H,W) or np.ndarray (H,W)/(H,W,C), grayscale or rgb.
        Returns List[Dict]: 'keypoints' (N,2) (x,y), 'scores' (N,), 'descriptors' (N,64); sorted by score, N <= top_k.
        """
        if top_k is None: top_k = self.top_k
        if detection_threshold is None: detection_threshold = self.detection_threshold
        out = self._detect_sparse_device(x, top_k, detection_threshold)
        n = out["n_valid"].tolist()                                   # the only host sync
        return [{"keypoints": out["keypoints"][b, :n[b]], "scores": out["scores"][b, :n[b]],
                 "descriptors": out["descriptors"][b, :n[b]]} for b in range(len(n))]

    def _mnn_device(self, f1, n1, n1_max, stride1, f2, n2, n2_max, stride2, batch, min_cossim, abs_bound=0.0):
        """abs_bound > 0: the caller guarantees max|f| <= abs_bound (1.0 for the sparse path's unit-norm descriptors), which
        spares the tensor-core matcher its max-reduction pass over both descriptor sets."""
        idx0 = self._empty((batch, n1_max), torch.int64)
        idx1 = self._empty((batch, n1_max), torch.int64)
        cnt = self._empty((batch,), torch.int32)
        ws = self._workspace(self._lib.xfeat_mnn_workspace_bytes(batch, n1_max, n2_max))
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_mnn_match_bounded(f1.data_ptr(), _ptr(n1), n1_max, stride1, f2.data_ptr(), _ptr(n2),
                                                         n2_max, stride2, batch, float(min_cossim), float(abs_bound),
                                                         idx0.data_ptr(), idx1.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                                                         ws.numel(), self._stream()),
                       "xfeat_mnn_match_bounded")
        return idx0, idx1, cnt

    @staticmethod
    def _as_desc(f: torch.Tensor, dev) -> torch.Tensor:
        f = f.to(dev)
        if f.dtype != torch.float32:
            f = f.float()
        return f.contiguous()

    @torch.inference_mode()
    def match(self, feats1, feats2, min_cossim=0.82):
        """Mutual nearest neighbours of two descriptor sets (reference: xfeat.py:327-348) -> (idx0, idx1) int64."""
        f1, f2 = self._as_desc(feats1, self.dev), self._as_desc(feats2, self.dev)
        if f1.shape[-1] != 64 or f2.shape[-1] != 64:
            raise RuntimeError("descriptors must be (N,64)")
        if len(f1) == 0 or len(f2) == 0:
            e = torch.empty((0,), dtype=torch.int64, device=self.dev)
            return e, e.clone()
        idx0, idx1, cnt = self._mnn_device(f1, None, len(f1), 0, f2, None, len(f2), 0, 1, min_cossim)
        n = int(cnt.item())
        return idx0[0, :n], idx1[0, :n]

    @torch.inference_mode()
    def batch_match(self, feats1, feats2, min_cossim=-1):
        """Batched MNN on raw dot products (reference: xfeat.py:265-290) -> list of B (idx0_b, idx1_b)."""
        f1, f2 = self._as_desc(feats1, self.dev), self._as_desc(feats2, self.dev)
        B, n1, _ = f1.shape
        n2 = f2.shape[1]
        idx0, idx1, cnt = self._mnn_device(f1, None, n1, n1 * 64, f2, None, n2, n2 * 64, B, min_cossim)
        c = cnt.tolist()
        return [(idx0[b, :c[b]], idx1[b, :c[b]]) for b in range(B)]

    def _match_sparse_batch_device(self, imgs1, imgs2, top_k: int, min_cossim: float, div255: bool = False):
        """Extraction of both image sets + per-pair MNN + keypoint gather, all on the device (no host sync).
        Returns mkpts0, mkpts1 (B,top_k,2) and n_matches (B) int32."""
        x1, x2 = self._to_bchw(imgs1), self._to_bchw(imgs2)
        B = x1.shape[0]
        if x1.shape == x2.shape and x1.dtype == x2.dtype:
            o = self._detect_sparse_device([x1, x2], top_k, self.detection_threshold, div255)
            k1, k2 = o["keypoints"][:B], o["keypoints"][B:]
            d1, d2 = o["descriptors"][:B], o["descriptors"][B:]
            n1, n2 = o["n_valid"][:B], o["n_valid"][B:]
        else:
            o1 = self._detect_sparse_device(x1, top_k, self.detection_threshold, div255)
            o2 = self._detect_sparse_device(x2, top_k, self.detection_threshold, div255)
            k1, k2, d1, d2, n1, n2 = (o1["keypoints"], o2["keypoints"], o1["descriptors"], o2["descriptors"],
                                      o1["n_valid"], o2["n_valid"])
        idx0, idx1, cnt = self._mnn_device(d1, n1, top_k, top_k * 64, d2, n2, top_k, top_k * 64, B, min_cossim,
                                           abs_bound=1.0)   # xfeat_detect_sparse writes unit-norm rows
        mk0, mk1 = self._empty((B, top_k, 2)), self._empty((B, top_k, 2))
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_gather_matches(k1.data_ptr(), k2.data_ptr(), top_k, top_k, idx0.data_ptr(),
                                                      idx1.data_ptr(), cnt.data_ptr(), B, mk0.data_ptr(), mk1.data_ptr(),
                                                      self._stream()), "xfeat_gather_matches")
        return mk0, mk1, cnt

    @torch.inference_mode()
    def match_xfeat(self, img1, img2, top_k=None, min_cossim=-1) -> Tuple[np.ndarray, np.ndarray]:
        """Extract + MNN match one image pair (reference: xfeat.py:165-186; only batch item 0 is used, as there).
        Returns mkpts_0, mkpts_1 -> np.ndarray (N,2)."""
        if top_k is None: top_k = self.top_k
        img1, d1 = self._parse_input(img1)
        img2, d2 = self._parse_input(img2)
        mk0, mk1, cnt = self._match_sparse_batch_device(img1[:1], img2[:1], top_k, min_cossim, div255=d1)
        n = int(cnt.item())
        return mk0[0, :n].cpu().numpy(), mk1[0, :n].cpu().numpy()

    @torch.inference_mode()
    def match_xfeat_batch(self, imgs1, imgs2, top_k=None, min_cossim=-1) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Batched extension of match_xfeat (the reference has none: xfeat.py:169): pair b = (imgs1[b], imgs2[b])."""
        if top_k is None: top_k = self.top_k
        imgs1, d1 = self._parse_input(imgs1)
        imgs2, _ = self._parse_input(imgs2)
        mk0, mk1, cnt = self._match_sparse_batch_device(imgs1, imgs2, top_k, min_cossim, div255=d1)
        c = cnt.tolist()
        mk0, mk1 = mk0.cpu().numpy(), mk1.cpu().numpy()
        return [(mk0[b, :c[b]], mk1[b, :c[b]]) for b in range(len(c))]

    # ------------------------------------------------------------------------------------------------------------
    # semi-dense path
    # ------------------------------------------------------------------------------------------------------------
    def _extract_dense_into(self, x: torch.Tensor, div255: bool, top_k: int, div_scale: float, scale_value: float,
                            out_rows: int, out_offset: int, kpts, desc, scales):
        """extractDense (xfeat.py:356-377) writing k rows at `out_offset`; returns k."""
        B, _, Hi, Wi = x.shape
        H, W = (Hi // 32) * 32, (Wi // 32) * 32
        rh, rw = Hi / H, Wi / W
        xn = self._preprocess(x, H, W, div255)
        feats, _, rel, _ = self._run_net(xn, B, H, W)
        cells = (H // 8) * (W // 8)
        k = min(cells, top_k)
        ws = self._workspace(self._lib.xfeat_dense_workspace_bytes(B, H, W, top_k))
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_detect_dense(self._ctx, feats.data_ptr(), rel.data_ptr(), B, H, W, top_k,
                                                    float(_F32(rw)), float(_F32(rh)), float(_F32(div_scale)),
                                                    float(_F32(scale_value)), out_rows, out_offset, kpts.data_ptr(),
                                                    desc.data_ptr(), _ptr(scales), None, ws.data_ptr(), ws.numel(),
                                                    self._stream()), "xfeat_detect_dense")
        return k

    def _resize(self, x: torch.Tensor, div255: bool, s: float) -> torch.Tensor:
        """F.interpolate(x, scale_factor=s, mode='bilinear', align_corners=False) (xfeat.py:380-381)."""
        x, code = self._img_args(x)
        B, C, Hi, Wi = x.shape
        Ho, Wo = int(math.floor(Hi * s)), int(math.floor(Wi * s))
        out = self._empty((B, C, Ho, Wo))
        sb, sc, sh, sw = x.stride()
        inv = float(_F32(1.0 / s))
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_resize_bilinear(x.data_ptr(), code, B, C, Hi, Wi, sb, sc, sh, sw, int(div255),
                                                       out.data_ptr(), Ho, Wo, inv, inv, self._stream()),
                       "xfeat_resize_bilinear")
        return out

    def _dense_device(self, x, top_k: int, multiscale: bool, div255: bool = False):
        x = self._to_bchw(x)
        B, _, Hi, Wi = x.shape

        def cells_of(h, w):
            return ((h // 32) * 32 // 8) * ((w // 32) * 32 // 8)

        if multiscale:
            s1, s2 = 0.6, 1.3                                                    # xfeat.py:379
            k1 = min(cells_of(int(math.floor(Hi * s1)), int(math.floor(Wi * s1))), int(top_k * 0.20))
            k2 = min(cells_of(int(math.floor(Hi * s2)), int(math.floor(Wi * s2))), int(top_k * 0.80))
            K = k1 + k2
            kpts, desc, scales = self._empty((B, K, 2)), self._empty((B, K, 64)), self._empty((B, K))
            x1 = self._resize(x, div255, s1)
            self._extract_dense_into(x1, False, int(top_k * 0.20), s1, 1 / s1, K, 0, kpts, desc, scales)
            del x1
            x2 = self._resize(x, div255, s2)
            self._extract_dense_into(x2, False, int(top_k * 0.80), s2, 1 / s2, K, k1, kpts, desc, scales)
        else:
            K = min(cells_of(Hi, Wi), top_k if top_k >= 1 else 100_000_000)
            kpts, desc, scales = self._empty((B, K, 2)), self._empty((B, K, 64)), self._empty((B, K))
            self._extract_dense_into(x, div255, K, 1.0, 1.0, K, 0, kpts, desc, scales)
        return {"keypoints": kpts, "descriptors": desc, "scales": scales}

    @torch.inference_mode()
    def detectAndComputeDense(self, x, top_k=None, multiscale=True) -> Dict[str, torch.Tensor]:
        """Dense *coarse* descriptors, batched (reference: xfeat.py:105-128): dict of 'keypoints' (B,K,2),
        'descriptors' (B,K,64) un-normalised, 'scales' (B,K); sorted by reliability per scale."""
        if top_k is None: top_k = self.top_k
        return self._dense_device(x, top_k, multiscale)

    def _refine_device(self, d0, d1, idx0, idx1, cnt, fine_conf: float = 0.25):
        B, K, _ = d0["descriptors"].shape
        matches = self._empty((B, K, 4))
        n_ref = self._empty((B,), torch.int32)
        ws = self._workspace(self._lib.xfeat_refine_workspace_bytes(B, K))
        with torch.cuda.device(self.dev):
            _lib.check(self._lib.xfeat_refine(self._ctx, d0["descriptors"].data_ptr(), d1["descriptors"].data_ptr(),
                                              d0["keypoints"].data_ptr(), d1["keypoints"].data_ptr(),
                                              d0["scales"].data_ptr(), idx0.data_ptr(), idx1.data_ptr(), cnt.data_ptr(),
                                              B, K, float(fine_conf), matches.data_ptr(), n_ref.data_ptr(),
                                              ws.data_ptr(), ws.numel(), self._stream()), "xfeat_refine")
        return matches, n_ref

    def _match_star_device(self, im_set1, im_set2, top_k: int, div255: bool = False):
        o1 = self._dense_device(im_set1, top_k, True, div255)
        o2 = self._dense_device(im_set2, top_k, True, div255)
        B, K, _ = o1["descriptors"].shape
        if o2["descriptors"].shape != o1["descriptors"].shape:
            raise RuntimeError("match_xfeat_star needs both image sets at the same resolution and batch size")
        idx0, idx1, cnt = self._mnn_device(o1["descriptors"], None, K, K * 64, o2["descriptors"], None, K, K * 64, B, -1)
        return self._refine_device(o1, o2, idx0, idx1, cnt)

    @torch.inference_mode()
    def match_xfeat_star(self, im_set1, im_set2, top_k=None):
        """Semi-dense matching with refinement, batched (reference: xfeat.py:188-217).
        B > 1 -> List[Tensor (N,4)] (x1,y1,x2,y2) on device; B == 1 -> two np.ndarray (N,2)."""
        if top_k is None: top_k = self.top_k
        im_set1, d1 = self._parse_input(im_set1)
        im_set2, _ = self._parse_input(im_set2)
        matches, n_ref = self._match_star_device(im_set1, im_set2, top_k, div255=d1)
        n = n_ref.tolist()
        out = [matches[b, :n[b]] for b in range(len(n))]
        return out if len(n) > 1 else (out[0][:, :2].cpu().numpy(), out[0][:, 2:].cpu().numpy())

    @torch.inference_mode()
    def match_lighterglue(self, d0, d1, min_conf=0.1):
        # reference: xfeat.py:131-162 -- needs kornia's LightGlue, which is outside this hot path
        raise RuntimeError('We rely on kornia for LightGlue. Install with: pip install kornia')

    # ------------------------------------------------------------------------------------------------------------
    # input helpers (reference semantics)
    # ------------------------------------------------------------------------------------------------------------
    def _parse_input(self, x):
        """parse_input (xfeat.py:396-403): 3-D -> add batch dim; numpy (B,H,W,C) uint8 -> the '/255' is applied
        on the device (flag), tensors pass through un-scaled.  Returns (tensor (B,C,H,W), div255 flag)."""
        if len(x.shape) == 3:
            x = x[None, ...]
        if isinstance(x, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
            if t.dtype != torch.uint8:
                return t.float() / 255, False
            return t, True
        return x, False

    def parse_input(self, x):
        """Reference-compatible parse_input (xfeat.py:396-403) returning a float tensor."""
        if len(x.shape) == 3:
            x = x[None, ...]
        if isinstance(x, np.ndarray):
            x = torch.tensor(x).permute(0, 3, 1, 2) / 255
        return x
