#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the LIVE, UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tools/make_golden.py

The reference has no tests or golden vectors of its own, so these files are
what pins ``oracle/xfeat_oracle.py`` to the reference's behaviour.  Inputs that cannot be regenerated
bit-identically elsewhere are stored inside the fixtures.
"""
import os
import sys

os.environ["CUDA_VISIBLE_DEVICES"] = ""
import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
import cv2  # noqa: E402
from modules.xfeat import XFeat  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.manual_seed(0)
xf = XFeat()  # default weights/xfeat.pt, top_k=4096, detection_threshold=0.05


def stage_probe(t: torch.Tensor, n: int = 2048, seed: int = 7):
    """A compact fingerprint of a stage tensor: sum, abs-sum and n fixed pseudo-random samples."""
    flat = t.detach().reshape(-1).double()
    rs = np.random.RandomState(seed)
    idx = rs.randint(0, flat.numel(), size=n)
    return dict(shape=np.array(t.shape), sum=np.float64(flat.sum()), asum=np.float64(flat.abs().sum()),
                idx=idx.astype(np.int64), val=t.detach().reshape(-1)[idx].numpy())


def save(name, **kw):
    flat = {}
    for k, v in kw.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}.{kk}"] = np.asarray(vv)
        else:
            flat[k] = np.asarray(v)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **flat)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(flat)} arrays")


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------
ref = cv2.imread(f"{REF}/assets/ref.png")
tgt = cv2.imread(f"{REF}/assets/tgt.png")
assert ref.shape == (600, 800, 3)
ref_vga = cv2.resize(ref, (640, 480))          # INTER_LINEAR; BASELINE config 1 "VGA pair"
tgt_vga = cv2.resize(tgt, (640, 480))
save("inputs_assets_vga.npz", ref=ref_vga, tgt=tgt_vga)

with torch.inference_mode():
    # -----------------------------------------------------------------------------------------
    # G1: sparse path on the VGA asset pair (numpy BGR u8 HWC -> parse_input /255), top_k=4096
    # -----------------------------------------------------------------------------------------
    x1 = xf.parse_input(ref_vga)
    x2 = xf.parse_input(tgt_vga)
    xx = torch.cat([x1, x2], 0)
    # stage fingerprints of the net for the 2-image batch
    xp, rh, rw = xf.preprocess_tensor(xx)
    M, K, Hh = xf.net(xp)
    heat = xf.get_kpts_heatmap(K)
    pos = xf.NMS(heat, threshold=0.05, kernel_size=5)
    n_cand = [(int(((heat[b] == torch.nn.functional.max_pool2d(heat[b:b + 1], 5, 1, 2)[0]) & (heat[b] > 0.05)).sum()))
              for b in range(2)]
    outs = xf.detectAndCompute(xx, top_k=4096)
    i0, i1 = xf.match(outs[0]["descriptors"], outs[1]["descriptors"], min_cossim=-1)
    j0, j1 = xf.match(outs[0]["descriptors"], outs[1]["descriptors"], min_cossim=0.82)
    mk0, mk1 = xf.match_xfeat(ref_vga, tgt_vga, top_k=4096)
    save("g1_sparse_vga.npz",
         feats=stage_probe(M), kpt_logits=stage_probe(K), reliability=stage_probe(Hh), heat=stage_probe(heat),
         nms_pos=pos.numpy(), n_cand=np.array(n_cand),
         kp0=outs[0]["keypoints"].numpy(), sc0=outs[0]["scores"].numpy(), desc0=outs[0]["descriptors"].numpy(),
         kp1=outs[1]["keypoints"].numpy(), sc1=outs[1]["scores"].numpy(),
         desc1_probe=stage_probe(outs[1]["descriptors"]),
         match_idx0=i0.numpy(), match_idx1=i1.numpy(), match082_idx0=j0.numpy(), match082_idx1=j1.numpy(),
         mkpts0=mk0, mkpts1=mk1)
    print("G1: cand", n_cand, "kpts", [len(o["keypoints"]) for o in outs], "matches", len(i0), len(j0), mk0.shape)

    # -----------------------------------------------------------------------------------------
    # G2: non-/32 size (crop 300x400 of the VGA pair -> internal 288x384), float tensor input path,
    #     small top_k so the cut is exercised; B=2
    # -----------------------------------------------------------------------------------------
    crop = np.stack([ref_vga[100:400, 120:520], tgt_vga[100:400, 120:520]])            # (2,300,400,3) u8
    xc = torch.tensor(crop).permute(0, 3, 1, 2).float()                                    # un-scaled 0..255 floats
    outs = xf.detectAndCompute(xc, top_k=500)
    xp, rh, rw = xf.preprocess_tensor(xc)
    M, K, Hh = xf.net(xp)
    save("g2_sparse_crop.npz", rh=rh, rw=rw, xp=stage_probe(xp), feats=stage_probe(M), kpt_logits=stage_probe(K),
         kp0=outs[0]["keypoints"].numpy(), sc0=outs[0]["scores"].numpy(), desc0=outs[0]["descriptors"].numpy(),
         kp1=outs[1]["keypoints"].numpy(), sc1=outs[1]["scores"].numpy(), desc1=outs[1]["descriptors"].numpy())
    print("G2: kpts", [len(o["keypoints"]) for o in outs], rh, rw)

    # -----------------------------------------------------------------------------------------
    # G3: stored small randn input (B=2, 3x96x128), every backbone stage fingerprinted
    # -----------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    xr = torch.randn(2, 3, 96, 128, generator=g)
    hooks, acts = [], {}
    net = xf.net
    names = {"block1": net.block1, "block2": net.block2, "block3": net.block3, "block4": net.block4,
             "block5": net.block5, "block_fusion": net.block_fusion, "skip1": net.skip1, "norm": net.norm}
    for n, m in names.items():
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.detach().clone())))
    M, K, Hh = net(xr)
    for h in hooks:
        h.remove()
    outs = xf.detectAndCompute(xr, top_k=256)
    save("g3_randn_small.npz", x=xr.numpy(), feats=M.numpy(), kpt_logits=K.numpy(), reliability=Hh.numpy(),
         **{f"act_{n}": stage_probe(a) for n, a in acts.items()},
         kp0=outs[0]["keypoints"].numpy(), sc0=outs[0]["scores"].numpy(), desc0=outs[0]["descriptors"].numpy(),
         kp1=outs[1]["keypoints"].numpy(), sc1=outs[1]["scores"].numpy(), desc1=outs[1]["descriptors"].numpy())
    print("G3: kpts", [len(o["keypoints"]) for o in outs])

    # -----------------------------------------------------------------------------------------
    # G4: semi-dense path (match_xfeat_star) on the VGA asset pair, B=2 (pair and swapped pair) and B=1
    # -----------------------------------------------------------------------------------------
    s1 = torch.cat([x1, x2], 0)
    s2 = torch.cat([x2, x1], 0)
    d1 = xf.detectAndComputeDense(s1, top_k=4096)
    d2 = xf.detectAndComputeDense(s2, top_k=4096)
    idxs = xf.batch_match(d1["descriptors"], d2["descriptors"])
    ml = xf.match_xfeat_star(s1, s2, top_k=4096)
    a0, a1 = xf.match_xfeat_star(ref_vga, tgt_vga, top_k=4096)
    save("g4_star_vga.npz",
         kp=d1["keypoints"].numpy(), scales=d1["scales"].numpy(), desc_probe=stage_probe(d1["descriptors"]),
         desc_b0_head=d1["descriptors"][0, :256].numpy(),
         coarse0_idx0=idxs[0][0].numpy(), coarse0_idx1=idxs[0][1].numpy(),
         coarse1_idx0=idxs[1][0].numpy(), coarse1_idx1=idxs[1][1].numpy(),
         matches0=ml[0].numpy(), matches1=ml[1].numpy(), b1_mk0=a0, b1_mk1=a1)
    print("G4: coarse", [len(i[0]) for i in idxs], "refined", [len(m) for m in ml], a0.shape)

    # -----------------------------------------------------------------------------------------
    # G5: MNN on stored unit-norm descriptors (N1=700, N2=512) incl. exact-duplicate rows (ties)
    # -----------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    f1 = torch.nn.functional.normalize(torch.randn(700, 64, generator=g), dim=-1)
    f2 = torch.nn.functional.normalize(torch.randn(512, 64, generator=g), dim=-1)
    f2[100] = f2[7]          # duplicated column descriptor: row argmax tie -> first index
    f1[300] = f1[20]         # duplicated row descriptor: column argmax tie -> first index
    f1[650] = f2[33]         # a perfect match (cos = 1)
    a0_, a1_ = xf.match(f1, f2, min_cossim=-1)
    b0_, b1_ = xf.match(f1, f2, min_cossim=0.82)
    c0_, c1_ = xf.match(f1, f2, min_cossim=0.3)
    save("g5_mnn.npz", f1=f1.numpy(), f2=f2.numpy(), idx0=a0_.numpy(), idx1=a1_.numpy(),
         idx0_082=b0_.numpy(), idx1_082=b1_.numpy(), idx0_03=c0_.numpy(), idx1_03=c1_.numpy())
    print("G5: matches", len(a0_), len(b0_), len(c0_))
