"""GPU parity, end to end through the drop-in XFeat class (own backbone -> own selection -> own matcher) against the
oracle on the same inputs, plus API-shape conformance with the reference's minimal_example.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import xfeat_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def xf():
    from accelerated_features_b200 import XFeat
    return XFeat()


from tests.parity_util import (assert_keypoints_exact_modulo_ties, assert_matches_exact_modulo_ties, kp_set, match_set,  # noqa: E402
                               record)


def _check_detect(name, got, want, st, top_k, rw=1.0, rh=1.0):
    """detectAndCompute output of one batch vs the oracle: exact keypoint sets modulo enumerated near-ties, scores and
    descriptors within 1e-3 (north_star) on the common keypoints.  Returns per-image (n_got, n_want, n_diff)."""
    stats = []
    for b in range(len(want)):
        g, w = got[b], want[b]
        gk, wk = g["keypoints"].cpu().numpy(), w["keypoints"].numpy()
        nd = assert_keypoints_exact_modulo_ties(f"{name}[{b}]", gk, wk, st, b, top_k, rw, rh)
        gi = {(float(x), float(y)): i for i, (x, y) in enumerate(gk)}
        wi = {(float(x), float(y)): i for i, (x, y) in enumerate(wk)}
        common = sorted(set(gi) & set(wi))
        derr = serr = 0.0
        if common:
            ia = np.array([gi[c] for c in common]); ib = np.array([wi[c] for c in common])
            derr = float(np.abs(g["descriptors"].cpu().numpy()[ia] - w["descriptors"].numpy()[ib]).max())
            serr = float(np.abs(g["scores"].cpu().numpy()[ia] - w["scores"].numpy()[ib]).max())
        assert derr < 1e-3 and serr < 1e-3, (name, b, derr, serr)
        s = g["scores"].cpu().numpy()
        assert np.all(s[:-1] >= s[1:]) and np.all(s > 0)
        stats.append({"got": len(gk), "want": len(wk), "differing": nd, "desc_err": derr, "score_err": serr})
    return stats


def test_detect_and_compute_assets(xf, oracle_state, assets_vga):
    ref, tgt = assets_vga
    x = torch.cat([orc.parse_input(ref), orc.parse_input(tgt)], 0)
    want, st = orc.detect_and_compute(oracle_state, x, 4096, return_stages=True)
    got = xf.detectAndCompute(x, top_k=4096)
    assert len(got) == 2
    for g in got:
        assert g["keypoints"].dtype == torch.float32 and g["keypoints"].shape[1] == 2
        assert g["descriptors"].shape[1] == 64 and g["scores"].ndim == 1
    stats = _check_detect("assets_vga", got, want, st, 4096)
    record("detectAndCompute_assets_vga", images=stats)


def test_match_xfeat_assets_numpy_input(xf, oracle_state, assets_vga, golden):
    ref, tgt = assets_vga
    mk0, mk1 = xf.match_xfeat(ref, tgt, top_k=4096)           # numpy HWC uint8 -> /255 on device
    assert isinstance(mk0, np.ndarray) and mk0.dtype == np.float32 and mk0.shape == mk1.shape and mk0.shape[1] == 2
    g = golden("g1_sparse_vga.npz")                            # produced by the live reference (tools/make_golden.py)
    want = match_set(g["mkpts0"], g["mkpts1"])
    got = match_set(mk0, mk1)
    x = torch.cat([orc.parse_input(ref), orc.parse_input(tgt)], 0)
    res = orc.detect_and_compute(oracle_state, x, 4096)
    nd = assert_matches_exact_modulo_ties("match_xfeat_assets", got, want, res[0]["keypoints"].numpy(), res[0]["descriptors"],
                                          res[1]["keypoints"].numpy(), res[1]["descriptors"])
    record("match_xfeat_assets_vga", matches=len(got), golden=len(want), differing=nd)


def test_randn_vga_vs_oracle(xf, oracle_state):
    """The bench workload (BASELINE config 2: seed-0 randn VGA, top_k 4096, saturated cut) on 4 pairs, against the oracle:
    keypoints exact modulo enumerated near-ties (the 4096-cut included), matches exact modulo MNN near-ties."""
    g = torch.Generator().manual_seed(0)
    x1 = torch.randn(4, 3, 480, 640, generator=g)
    x2 = torch.randn(4, 3, 480, 640, generator=g)
    with torch.inference_mode():
        w1, st1 = orc.detect_and_compute(oracle_state, x1, 4096, return_stages=True)
        w2, st2 = orc.detect_and_compute(oracle_state, x2, 4096, return_stages=True)
    g1 = xf.detectAndCompute(x1, top_k=4096)
    g2 = xf.detectAndCompute(x2, top_k=4096)
    s1 = _check_detect("randn_vga_set1", g1, w1, st1, 4096)
    s2 = _check_detect("randn_vga_set2", g2, w2, st2, 4096)
    out = xf.match_xfeat_batch(x1, x2, top_k=4096)
    pairs = []
    for b in range(4):
        i0, i1 = orc.mnn_match(w1[b]["descriptors"], w2[b]["descriptors"], -1)
        want = match_set(w1[b]["keypoints"][i0].numpy(), w2[b]["keypoints"][i1].numpy())
        got = match_set(*out[b])
        nd = assert_matches_exact_modulo_ties(f"randn_vga_pair{b}", got, want, w1[b]["keypoints"].numpy(), w1[b]["descriptors"],
                                              w2[b]["keypoints"].numpy(), w2[b]["descriptors"], s1[b]["differing"],
                                              s2[b]["differing"])
        pairs.append({"matches": len(got), "oracle": len(want), "differing": nd})
    record("randn_vga_bench_workload", set1=s1, set2=s2, pairs=pairs)


def test_match_xfeat_batch_equals_single(xf, assets_vga):
    ref, tgt = assets_vga
    b1 = np.stack([ref, tgt, ref]); b2 = np.stack([tgt, ref, ref])
    out = xf.match_xfeat_batch(b1, b2, top_k=2048)
    assert len(out) == 3
    s0 = xf.match_xfeat(ref, tgt, top_k=2048)
    assert np.array_equal(out[0][0], s0[0]) and np.array_equal(out[0][1], s0[1])
    # identical images: every keypoint matches itself
    assert np.array_equal(out[2][0], out[2][1]) and len(out[2][0]) > 1500


def test_star_assets(xf, oracle_state, assets_vga, golden):
    ref, tgt = assets_vga
    g = golden("g4_star_vga.npz")
    x1, x2 = orc.parse_input(ref), orc.parse_input(tgt)
    s1, s2 = torch.cat([x1, x2], 0), torch.cat([x2, x1], 0)
    d = xf.detectAndComputeDense(s1, top_k=4096)
    assert d["keypoints"].shape == (2, 4095, 2) and d["descriptors"].shape == (2, 4095, 64) and d["scales"].shape == (2, 4095)
    assert np.array_equal(d["scales"].cpu().numpy(), g["scales"])
    for b in range(2):
        common = kp_set(d["keypoints"][b].cpu().numpy()) & kp_set(g["kp"][b])
        assert len(common) >= 4095 - 4, len(common)          # only reliability near-ties at the two top-k cuts may move
    ml = xf.match_xfeat_star(s1, s2, top_k=4096)
    assert isinstance(ml, list) and len(ml) == 2 and ml[0].shape[1] == 4 and ml[0].is_cuda
    for b in range(2):
        want = g[f"matches{b}"]
        got = ml[b].cpu().numpy()
        # compare as sets of (x2,y2) targets (exact cell coords) with refined sources within 0.05 px
        wd = {(float(r[2]), float(r[3])): r[:2] for r in want}
        hit = sum(1 for r in got if (float(r[2]), float(r[3])) in wd and np.abs(wd[(float(r[2]), float(r[3]))] - r[:2]).max() < 0.05)
        print(f"star pair {b}: {len(got)} vs {len(want)} refined, agreeing {hit}")
        record(f"match_xfeat_star_assets_vga_pair{b}", refined=len(got), golden=len(want), agreeing=hit)
        # coarse MNN near-ties and conf ~ 0.25 rows are the only admitted differences: a handful of rows at most
        assert len(want) - hit <= 6 and abs(len(got) - len(want)) <= 6
    a0, a1 = xf.match_xfeat_star(ref, tgt, top_k=4096)         # B == 1 -> numpy pair
    assert isinstance(a0, np.ndarray) and a0.shape == a1.shape and a0.shape[1] == 2


def test_constant_and_featureless_images(xf, oracle_state):
    """Images without structure: all-zero and constant images give NO keypoints (InstanceNorm of a constant is 0), a horizontal
    ramp a few dozen; the batch mixes them with a random image, and a pair with an empty side yields no matches (the reference's
    single-pair `match` raises IndexError from torch.max there, xfeat.py:331-335; a batch must not abort)."""
    g = torch.Generator().manual_seed(5)
    ramp = torch.linspace(0, 1, 128).view(1, 1, 1, 128).expand(1, 3, 96, 128).contiguous()
    x = torch.cat([torch.zeros(1, 3, 96, 128), torch.full((1, 3, 96, 128), 0.37), ramp, torch.randn(1, 3, 96, 128, generator=g)])
    want, st = orc.detect_and_compute(oracle_state, x, 512, return_stages=True)
    got = xf.detectAndCompute(x, top_k=512)
    assert [len(w["keypoints"]) for w in want][:2] == [0, 0] and len(want[2]["keypoints"]) > 0
    stats = _check_detect("featureless", got, want, st, 512)
    assert [s["got"] for s in stats] == [s["want"] for s in stats]
    for b in (0, 1):
        assert got[b]["keypoints"].shape == (0, 2) and got[b]["descriptors"].shape == (0, 64) and got[b]["scores"].shape == (0,)
    pairs = xf.match_xfeat_batch(x, x.flip(0), top_k=512)          # (zeros,randn) (const,ramp) (ramp,const) (randn,zeros)
    assert all(a.shape == (0, 2) and b.shape == (0, 2) for a, b in pairs)
    i0, i1 = xf.match(got[0]["descriptors"], got[3]["descriptors"])
    assert i0.numel() == 0 and i1.numel() == 0
    same = xf.match_xfeat_batch(x, x, top_k=512)                   # identical sets: a random image's keypoints match themselves
    assert len(same[3][0]) == stats[3]["got"] and np.array_equal(same[3][0], same[3][1])
    assert all(len(a) <= s["got"] for (a, _), s in zip(same, stats))   # (the ramp repeats descriptors down its columns: ties)
    record("featureless_images", keypoints=[s["got"] for s in stats], differing=sum(s["differing"] for s in stats))


def test_run_to_run_bitwise_reproducible(xf):
    """The sparse VGA path has no order-dependent arithmetic (the image statistics are reduced in rank order by the cluster
    kernel, candidate lists are sorted by unique keys, arg-max ties go to the lowest index): repeated runs give the same bits."""
    g = torch.Generator().manual_seed(11)
    x1 = torch.randn(8, 3, 480, 640, generator=g).cuda(); x2 = torch.randn(8, 3, 480, 640, generator=g).cuda()
    runs = []
    for _ in range(4):
        mk0, mk1, cnt = xf._match_sparse_batch_device(x1, x2, 4096, -1)
        c = cnt.tolist()
        runs.append((c, [mk0[b, :c[b]].clone() for b in range(8)], [mk1[b, :c[b]].clone() for b in range(8)]))
    for c, a0, a1 in runs[1:]:
        assert c == runs[0][0]
        assert all(torch.equal(u, v) for u, v in zip(a0, runs[0][1])) and all(torch.equal(u, v) for u, v in zip(a1, runs[0][2]))
    o = [xf._detect_sparse_device(x1, 4096, 0.05) for _ in range(3)]
    for k in ("keypoints", "scores", "descriptors", "n_valid"):
        assert torch.equal(o[0][k], o[1][k]) and torch.equal(o[0][k], o[2][k]), k


def test_minimal_example_api(xf):
    """the reference's minimal_example.py sequence (shapes only; randn inputs)."""
    torch.manual_seed(0)
    x = torch.randn(1, 3, 480, 640)
    out = xf.detectAndCompute(x, top_k=4096)[0]
    assert out["keypoints"].shape[1] == 2 and out["descriptors"].shape[1] == 64
    assert out["keypoints"].shape[0] == out["scores"].shape[0] == out["descriptors"].shape[0] <= 4096
    outs = xf.detectAndCompute(torch.randn(4, 3, 480, 640), top_k=4096)
    assert len(outs) == 4
    mk0, mk1 = xf.match_xfeat(torch.randn(1, 3, 480, 640), torch.randn(1, 3, 480, 640))
    assert mk0.shape == mk1.shape
    ml = xf.match_xfeat_star(torch.randn(4, 3, 480, 640), torch.randn(4, 3, 480, 640))
    assert len(ml) == 4 and ml[0].shape[1] == 4
    with pytest.raises(RuntimeError):
        xf.detectAndCompute(torch.randn(3, 480, 640))          # non-4D tensor (xfeat.py:230-231)
    with pytest.raises(RuntimeError):
        xf.detectAndCompute(np.zeros((2, 3, 4, 5), np.float32))  # bad numpy rank (xfeat.py:227)


def test_full_size_properties(xf):
    """BASELINE config 2 size (batch 64 VGA, top_k 4096): size-independent invariants."""
    g = torch.Generator().manual_seed(0)
    x1 = torch.randn(64, 3, 480, 640, generator=g)
    mk0, mk1, cnt = xf._match_sparse_batch_device(x1, x1, 4096, -1)      # identical sets: identity matching
    torch.cuda.synchronize()
    c = cnt.tolist()
    assert min(c) > 3000
    for b in (0, 17, 63):
        assert torch.equal(mk0[b, :c[b]], mk1[b, :c[b]])
    x2 = torch.randn(64, 3, 480, 640, generator=g)
    o = xf._detect_sparse_device(torch.cat([x1[:8], x2[:8]]), 4096, 0.05)
    d = o["descriptors"]; n = o["n_valid"].tolist()
    # B >= 32 runs the multi-CTA-per-image spatially ordered sampler, B = 8 the generic one: same arithmetic, same bits
    o32 = xf._detect_sparse_device(torch.cat([x1[:16], x2[:16]]), 4096, 0.05)
    o8 = xf._detect_sparse_device(x2[:8], 4096, 0.05)
    for k in ("keypoints", "scores", "descriptors", "n_valid"):
        assert torch.equal(o32[k][16:24], o8[k]), k
    nrm = d[0, :n[0]].norm(dim=-1)
    assert float((nrm - 1).abs().max()) < 1e-5                         # unit descriptors
    s = o["scores"][0, :n[0]]
    assert bool((s[:-1] >= s[1:]).all())                                # sorted by score
    # mutual matches are one-to-one
    idx0, idx1, cnt = xf._mnn_device(d[:8], o["n_valid"][:8], 4096, 4096 * 64, d[8:], o["n_valid"][8:], 4096, 4096 * 64, 8, -1)
    for b in range(8):
        m = int(cnt[b]); a = idx0[b, :m].cpu().numpy(); bb = idx1[b, :m].cpu().numpy()
        assert len(set(a)) == m and len(set(bb)) == m and np.all(np.diff(a) > 0)


def test_batch32_small_images_vs_oracle(xf, oracle_state):
    """B = 32 routes description through the one-CTA-per-image (spatially ordered) sampler: check it against the oracle."""
    g = torch.Generator().manual_seed(21)
    x = torch.randn(32, 3, 96, 128, generator=g)
    want, st = orc.detect_and_compute(oracle_state, x, 300, return_stages=True)
    got = xf.detectAndCompute(x, top_k=300)
    stats = _check_detect("batch32_small", got, want, st, 300)
    record("batch32_small_images", differing=sum(s["differing"] for s in stats), worst_desc_err=max(s["desc_err"] for s in stats))


def test_star_hd_pair_vs_oracle(xf, oracle_state, assets_vga):
    """BASELINE config 3 geometry (1280x960, dual scale 768x576 + 1664x1248) on structured images, B = 2, against the oracle."""
    ref, tgt = assets_vga
    x1 = torch.nn.functional.interpolate(orc.parse_input(ref), size=(960, 1280), mode="bilinear", align_corners=False)
    x2 = torch.nn.functional.interpolate(orc.parse_input(tgt), size=(960, 1280), mode="bilinear", align_corners=False)
    s1, s2 = torch.cat([x1, x2], 0), torch.cat([x2, x1], 0)
    with torch.inference_mode():
        want = orc.match_xfeat_star(oracle_state, s1, s2, 4096)
    got = xf.match_xfeat_star(s1, s2, top_k=4096)
    assert len(got) == 2
    for b in range(2):
        w, g = want[b].numpy(), got[b].cpu().numpy()
        wd = {(float(r[2]), float(r[3])): r[:2] for r in w}
        hit = sum(1 for r in g if (float(r[2]), float(r[3])) in wd and np.abs(wd[(float(r[2]), float(r[3]))] - r[:2]).max() < 0.05)
        print(f"HD star pair {b}: {len(g)} vs {len(w)} refined matches, agreeing {hit}")
        record(f"match_xfeat_star_hd_pair{b}", refined=len(g), oracle=len(w), agreeing=hit)
        assert len(w) - hit <= 8 and abs(len(g) - len(w)) <= 8


@pytest.mark.parametrize("hw", [(32, 32), (64, 96), (600, 800), (200, 328)])
def test_odd_sizes_vs_oracle(xf, oracle_state, assets_vga, hw):
    """Geometry corner cases of the tensor-core tiles: 1x1 maps at 1/32, partial tiles, non-/32 inputs (resize path)."""
    ref, tgt = assets_vga
    H, W = hw
    x = torch.nn.functional.interpolate(torch.cat([orc.parse_input(ref), orc.parse_input(tgt)], 0), size=(H, W), mode="bilinear",
                                        align_corners=False)
    want, st = orc.detect_and_compute(oracle_state, x, 512, return_stages=True)
    got = xf.detectAndCompute(x, top_k=512)
    stats = _check_detect(f"odd_{H}x{W}", got, want, st, 512, rw=st["rw"], rh=st["rh"])
    record(f"odd_size_{H}x{W}", images=stats)
