"""Readers for tests/golden/*.npz (written by oracle/gen_golden.py from the reference)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def numpy_pinned():
    """True when NumPy's float32 exp/log/tanh are the libm ones the goldens were made with."""
    return os.environ.get("MGD_NUMPY_PIN_EFFECTIVE") == "1"


def files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


def anchors_of(z):
    dt = np.float64 if bool(z["anchors_f64"]) else np.float32
    return [np.array(a, dtype=dt) for a in z["anchors"]]


def dense_y_true(z, prefix=""):
    out = []
    l = 0
    while f"{prefix}shape{l}" in z:
        y = np.zeros(tuple(z[f"{prefix}shape{l}"]), dtype=np.float32)
        idx = z[f"{prefix}idx{l}"]
        y[idx[:, 0], idx[:, 1], idx[:, 2]] = z[f"{prefix}val{l}"]
        out.append(y)
        l += 1
    return out


def preds_of(z):
    out = []
    l = 0
    while f"pred{l}" in z:
        out.append(z[f"pred{l}"].astype(np.float32))
        l += 1
    return out


def knobs_of(z):
    for k in range(int(z["n_knobs"])):
        yield k, dict(image_shape=tuple(int(v) for v in z[f"k{k}_image_shape"]),
                      confidence=float(z[f"k{k}_conf"]), nms_threshold=float(z[f"k{k}_thr"]),
                      nms_method=str(z[f"k{k}_method"]), max_boxes=int(z[f"k{k}_max"]))


def assert_encode_matches(got, ref, exact_floats):
    for g, r in zip(got, ref):
        g = np.asarray(g)
        assert g.shape == r.shape and g.dtype == np.float32
        assert np.array_equal(g[..., 4:], r[..., 4:]), "mask / one-hot channels differ"
        assert np.array_equal(g[..., 0:2], r[..., 0:2]), "cell offsets differ"
        if exact_floats:
            assert np.array_equal(g[..., 2:4], r[..., 2:4])
        else:
            np.testing.assert_allclose(g[..., 2:4], r[..., 2:4], rtol=1e-5, atol=1e-6)


def coco608_inputs(c_oracle_encode):
    """Regenerate the inputs of coco608_detections.npz from their seeds (oracle/gen_golden.py:
    coco608_inputs / coco608_case) and check them against the stored SHA-256."""
    import hashlib
    import torch
    from multigriddet_b200 import synth
    S, C = 608, 80
    anchors = synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(400, 2, 100, S, C, anchors=anchors)
    y = c_oracle_encode(boxes, (S, S), anchors, C)
    preds = [p.numpy() for p in synth.planted_head_outputs([torch.from_numpy(a) for a in y], 3, seed=401)]
    h = hashlib.sha256()
    for p in preds:
        h.update(np.ascontiguousarray(p).tobytes())
    return S, C, anchors, preds, h.hexdigest()
