"""Pin oracle/xfeat_oracle.py (the CPU restatement) to fixtures produced by the live reference
(tools/make_golden.py).  Integer outputs must be bit-exact; floats are produced by the same ATen ops
in the same order, tested at 1e-6 (observed: 0 ulp in the build container)."""
import numpy as np
import pytest
import torch

from oracle import xfeat_oracle as orc

ATOL = 1e-6


def probe_ok(t, g, key, atol=ATOL):
    t = t.detach()
    assert list(t.shape) == list(g[key + ".shape"]), key
    flat = t.reshape(-1)
    np.testing.assert_allclose(flat[torch.from_numpy(g[key + ".idx"])].numpy(), g[key + ".val"], atol=atol, rtol=1e-5,
                               err_msg=key)
    s = float(flat.double().sum())
    assert abs(s - float(g[key + ".sum"])) <= 1e-6 * max(1.0, float(g[key + ".asum"])), key


def test_g1_sparse_vga(golden, oracle_state, assets_vga):
    g = golden("g1_sparse_vga.npz")
    ref, tgt = assets_vga
    xx = torch.cat([orc.parse_input(ref), orc.parse_input(tgt)], 0)
    with torch.inference_mode():
        res, st = orc.detect_and_compute(oracle_state, xx, 4096, 0.05, return_stages=True)
    probe_ok(st["feats"], g, "feats")
    probe_ok(st["kpt_logits"], g, "kpt_logits")
    probe_ok(st["reliability"], g, "reliability")
    probe_ok(st["heat"], g, "heat")
    assert np.array_equal(st["nms_pos"].numpy(), g["nms_pos"])
    assert orc.nms_counts(st["heat"]) == list(g["n_cand"])
    for b in range(2):
        assert np.array_equal(res[b]["keypoints"].numpy(), g[f"kp{b}"])
        np.testing.assert_allclose(res[b]["scores"].numpy(), g[f"sc{b}"], atol=ATOL)
    np.testing.assert_allclose(res[0]["descriptors"].numpy(), g["desc0"], atol=ATOL)
    probe_ok(res[1]["descriptors"], g, "desc1_probe")
    i0, i1 = orc.mnn_match(res[0]["descriptors"], res[1]["descriptors"], -1)
    assert np.array_equal(i0.numpy(), g["match_idx0"]) and np.array_equal(i1.numpy(), g["match_idx1"])
    j0, j1 = orc.mnn_match(res[0]["descriptors"], res[1]["descriptors"], 0.82)
    assert np.array_equal(j0.numpy(), g["match082_idx0"]) and np.array_equal(j1.numpy(), g["match082_idx1"])
    with torch.inference_mode():
        mk0, mk1 = orc.match_xfeat(oracle_state, ref, tgt, 4096)
    assert np.array_equal(mk0, g["mkpts0"]) and np.array_equal(mk1, g["mkpts1"])
    # explicit numpy NMS rule == ATen NMS
    pos_np = orc.nms_numpy(st["heat"][0, 0].numpy(), 0.05)
    assert np.array_equal(pos_np, g["nms_pos"][0][: len(pos_np)])


def test_g2_crop_resize(golden, oracle_state, assets_vga):
    g = golden("g2_sparse_crop.npz")
    ref, tgt = assets_vga
    crop = np.stack([ref[100:400, 120:520], tgt[100:400, 120:520]])
    xc = torch.tensor(crop).permute(0, 3, 1, 2).float()
    with torch.inference_mode():
        res, st = orc.detect_and_compute(oracle_state, xc, 500, 0.05, return_stages=True)
    assert st["rh"] == float(g["rh"]) and st["rw"] == float(g["rw"])
    probe_ok(st["x"], g, "xp", atol=1e-4)
    probe_ok(st["feats"], g, "feats")
    for b in range(2):
        assert np.array_equal(res[b]["keypoints"].numpy(), g[f"kp{b}"])
        np.testing.assert_allclose(res[b]["scores"].numpy(), g[f"sc{b}"], atol=ATOL)
        np.testing.assert_allclose(res[b]["descriptors"].numpy(), g[f"desc{b}"], atol=ATOL)


def test_g3_randn_stages(golden, oracle_state):
    g = golden("g3_randn_small.npz")
    x = torch.from_numpy(g["x"])
    with torch.inference_mode():
        res, st = orc.detect_and_compute(oracle_state, x, 256, 0.05, return_stages=True)
    np.testing.assert_allclose(st["feats"].numpy(), g["feats"], atol=ATOL)
    np.testing.assert_allclose(st["kpt_logits"].numpy(), g["kpt_logits"], atol=ATOL, rtol=1e-5)
    np.testing.assert_allclose(st["reliability"].numpy(), g["reliability"], atol=ATOL)
    probe_ok(st["xn"], g, "act_norm")
    probe_ok(st["x1"], g, "act_block1")
    probe_ok(st["x2"], g, "act_block2")
    probe_ok(st["x3"], g, "act_block3")
    probe_ok(st["x4"], g, "act_block4")
    probe_ok(st["x5"], g, "act_block5")
    for b in range(2):
        assert np.array_equal(res[b]["keypoints"].numpy(), g[f"kp{b}"])
        np.testing.assert_allclose(res[b]["scores"].numpy(), g[f"sc{b}"], atol=ATOL)
        np.testing.assert_allclose(res[b]["descriptors"].numpy(), g[f"desc{b}"], atol=ATOL)


def test_g4_star(golden, oracle_state, assets_vga):
    g = golden("g4_star_vga.npz")
    ref, tgt = assets_vga
    x1, x2 = orc.parse_input(ref), orc.parse_input(tgt)
    s1, s2 = torch.cat([x1, x2], 0), torch.cat([x2, x1], 0)
    with torch.inference_mode():
        d1 = orc.detect_and_compute_dense(oracle_state, s1, 4096)
        d2 = orc.detect_and_compute_dense(oracle_state, s2, 4096)
        idxs = orc.batch_match(d1["descriptors"], d2["descriptors"])
        ml = orc.match_xfeat_star(oracle_state, s1, s2, 4096)
        a0, a1 = orc.match_xfeat_star(oracle_state, ref, tgt, 4096)
    assert np.array_equal(d1["keypoints"].numpy(), g["kp"])
    assert np.array_equal(d1["scales"].numpy(), g["scales"])
    probe_ok(d1["descriptors"], g, "desc_probe", atol=1e-5)
    np.testing.assert_allclose(d1["descriptors"][0, :256].numpy(), g["desc_b0_head"], atol=1e-5)
    for b in range(2):
        assert np.array_equal(idxs[b][0].numpy(), g[f"coarse{b}_idx0"])
        assert np.array_equal(idxs[b][1].numpy(), g[f"coarse{b}_idx1"])
        np.testing.assert_allclose(ml[b].numpy(), g[f"matches{b}"], atol=1e-4)
    np.testing.assert_allclose(a0, g["b1_mk0"], atol=1e-4)
    np.testing.assert_allclose(a1, g["b1_mk1"], atol=1e-4)


def test_g5_mnn(golden):
    g = golden("g5_mnn.npz")
    f1, f2 = torch.from_numpy(g["f1"]), torch.from_numpy(g["f2"])
    for thr, sfx in ((-1, ""), (0.82, "_082"), (0.3, "_03")):
        i0, i1 = orc.mnn_match(f1, f2, thr)
        assert np.array_equal(i0.numpy(), g["idx0" + sfx]) and np.array_equal(i1.numpy(), g["idx1" + sfx])
    # explicit numpy rule agrees, including first-index tie-breaks on the duplicated rows/cols
    s = (f1 @ f2.t()).numpy()
    n0, n1 = orc.mnn_numpy(s, -1)
    assert np.array_equal(n0, g["idx0"]) and np.array_equal(n1, g["idx1"])
    # batch_match on the same data (B=1) gives the same pairs
    b = orc.batch_match(f1[None, :512], f2[None])
    r0, r1 = orc.mnn_match(f1[:512], f2, -1)
    assert np.array_equal(b[0][0].numpy(), r0.numpy()) and np.array_equal(b[0][1].numpy(), r1.numpy())


def test_source_coord_convention():
    """ix = p*Wm/(W-1) - 0.5; nearest on the full-res map returns p except the last
    row/col, which rounds out of bounds (quirk A)."""
    for W in (640, 480, 384, 288, 128, 96):
        p = np.arange(W)
        ix = orc.source_coord(p, W, W)
        r = np.rint(ix).astype(np.int64)
        assert np.array_equal(r[:-1], p[:-1]) and r[-1] == W
        ixm = orc.source_coord(p, W, W // 8)
        np.testing.assert_allclose(ixm, p * (W // 8) / (W - 1) - 0.5, atol=2e-5)
