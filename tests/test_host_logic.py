"""Host-side logic that needs no GPU: weight folding / packing against the oracle's un-folded layers, layer table
consistency with the C header, input parsing rules."""
import os
import re

import numpy as np
import torch
import torch.nn.functional as F

from accelerated_features_b200 import weights as W
from oracle import xfeat_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_layer_table():
    src = open(os.path.join(ROOT, "accelerated_features_b200", "csrc", "layers.h")).read()
    body = src[src.index("kLayers[L_COUNT] = {"):]
    body = body[: body.index("};")]
    return [tuple(int(v) for v in m) for m in re.findall(r"\{\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\}", body)]


def test_layer_table_matches_state_dict(oracle_state):
    table = c_layer_table()
    assert len(table) == len(W.LAYERS) == 32
    for (cin, cout, ks, stride, relu), (conv, bn) in zip(table, W.LAYERS):
        w = oracle_state[conv + ".weight"]
        assert w.shape[0] == cout and w.shape[1] == cin
        assert (w.shape[2] if w.dim() == 4 else 1) == ks
        assert relu == (1 if bn is not None else 0)      # every BasicLayer / BN'd linear is followed by ReLU


def test_folded_layers_equal_conv_bn(oracle_state):
    sd = {k: v.numpy() for k, v in oracle_state.items()}
    g = torch.Generator().manual_seed(0)
    for conv, bn in [W.LAYERS[5], W.LAYERS[8], W.LAYERS[16], W.LAYERS[19], W.LAYERS[26]]:
        wk, b = W.fold_layer(sd, conv, bn)
        w = oracle_state[conv + ".weight"]
        cout, cin, ks = w.shape[0], w.shape[1], w.shape[2]
        x = torch.randn(1, cin, 9, 11, generator=g)
        if bn is not None:
            want = orc._basic_layer(oracle_state, conv[: -len(".layer.0")], x, 1, ks // 2)
        else:
            want = F.conv2d(x, w, oracle_state[conv + ".bias"], padding=ks // 2)
        wf = torch.from_numpy(wk).reshape(ks, ks, cin, cout).permute(3, 2, 0, 1).contiguous()
        got = F.conv2d(x, wf, torch.from_numpy(b), padding=ks // 2)
        if bn is not None:
            got = F.relu(got)
        assert (got - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())


def test_fine_matcher_fold(oracle_state):
    sd = {k: v.numpy() for k, v in oracle_state.items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(7, 128, generator=g)
    want = orc.fine_matcher(oracle_state, x)
    h = x
    for i, (lin, bn) in enumerate(W.LAYERS[27:]):
        wk, b = W.fold_layer(sd, lin, bn)
        h = h @ torch.from_numpy(wk) + torch.from_numpy(b)
        if bn is not None:
            h = F.relu(h)
    assert (h - want).abs().max().item() < 1e-4


def test_pack_layout_offsets():
    sd = W.random_state_dict(3)
    blob = W.pack_weights(sd)
    off = 0
    for (cin, cout, ks, _, _), (conv, bn) in zip(c_layer_table(), W.LAYERS):
        wk, b = W.fold_layer(sd, conv, bn)
        n = ks * ks * cin * cout
        assert np.array_equal(blob[off:off + n], wk.reshape(-1))
        off += (n + 3) // 4 * 4
        assert np.array_equal(blob[off:off + cout], b)
        off += (cout + 3) // 4 * 4
    assert off == blob.size


def test_unfold_index_rule():
    """channel 8i+j of the unfolded tensor is pixel (8h+i, 8w+j) -- the address rule of the IN_UNFOLD8 loader."""
    x = torch.arange(2 * 16 * 24, dtype=torch.float32).reshape(2, 1, 16, 24)
    u = orc.unfold8(x)
    for (b, i, j, h, w) in [(0, 0, 0, 0, 0), (1, 3, 5, 1, 2), (0, 7, 7, 1, 0)]:
        assert u[b, 8 * i + j, h, w] == x[b, 0, 8 * h + i, 8 * w + j]


def test_heatmap_unshuffle_rule():
    k = torch.randn(1, 65, 3, 4)
    heat = orc.kpts_heatmap(k)
    p = F.softmax(k, 1)
    for (i, j, h, w) in [(0, 0, 0, 0), (2, 7, 1, 3), (7, 1, 2, 0)]:
        assert heat[0, 0, 8 * h + i, 8 * w + j] == p[0, 8 * i + j, h, w]


def test_eval_metrics_restatement():
    """evalharness restates megadepth1500.py's metric functions (the reference module imports poselib, absent here)."""
    from accelerated_features_b200 import evalharness as ev
    auc = ev.error_auc([1, 2, 30, 4, 50])
    assert abs(auc["auc@5"] - 0.4) < 1e-12 and abs(auc["auc@10"] - 0.5) < 1e-12 and abs(auc["auc@20"] - 0.55) < 1e-12
    T = np.eye(4)
    T[:3, 3] = [1.0, 0.0, 0.0]
    c, s = np.cos(np.deg2rad(10.0)), np.sin(np.deg2rad(10.0))
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    t_err, r_err = ev.relative_pose_error(T, R, np.array([-2.0, 0.0, 0.0]))      # sign of t is not observable
    assert abs(t_err) < 1e-6 and abs(r_err - 10.0) < 1e-6
    m = ev.compute_maa([{"t_err": 1.0, "R_err": 3.0}, {"t_err": 30.0, "R_err": 2.0}])
    assert m["mAcc@5"] == 0.5 and m["mAcc@20"] == 0.5
