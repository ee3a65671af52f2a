"""torch.hub entry mirroring the reference's hubconf.py (hubconf.py:5-15): XFeat(pretrained, top_k, detection_threshold).
There is no network here: `pretrained=True` loads the packaged weights instead of downloading them."""
dependencies = ["torch"]


def XFeat(pretrained=True, top_k=4096, detection_threshold=0.05):
    from accelerated_features_b200 import XFeat as _XFeat
    from accelerated_features_b200 import weights as _w
    weights = _w.load_state_dict(_w.DEFAULT_WEIGHTS) if pretrained else None
    return _XFeat(weights, top_k=top_k, detection_threshold=detection_threshold)
