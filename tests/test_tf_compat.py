"""The TensorFlow encoder's semantics (MGD_FLAG_TF_COMPAT): reference
multigriddet/data/generators.py:2696-3390.  TensorFlow is not installed here; the pin is the
reference function's own source executed over a NumPy stand-in for its tf.* ops
(oracle/tf_shim.py: tests/test_oracle_vs_reference.py live, tests/golden/tfencode_*.npz as
fixtures).  The CPU tests below additionally check the oracle restatement against hand-derived
values for exactly the points where the TF encoder differs from the NumPy one
and the one input the reference's own tests pin; the GPU tests check the CUDA path against
that oracle bit for bit.
"""
import numpy as np
import pytest

from multigriddet_b200 import synth
from oracle import mgd_oracle as O

S, C = 608, 80


def _anchors():
    return synth.coco_anchors(np.float32)


def _layer_of(y):
    return [int((t[0, ..., 4] > 0).sum()) for t in y]


def test_reference_9cell_box_alignment_case():
    """The one input the reference pins (tests/test_9cell_alignment.py:28-60): centre
    (311.999, 311.999), 100x80, class 0, C=1, small-first anchors -> nine cells whose
    stored offsets all decode to the same centre."""
    anchors = [np.array(a, np.float32) for a in
               ([[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]],
                [[116, 90], [156, 198], [373, 326]])]
    cx = cy = 311.999
    box = np.array([[[cx - 50, cy - 40, cx + 50, cy + 40, 0]]], np.float32)
    grids = [(19, 19), (38, 38), (76, 76)]
    y = O.encode_targets_tf_compat(box, (608, 608), anchors, 1, grids)
    assert _layer_of(y) == [0, 0, 9]
    t = y[2][0]
    rows, cols = np.nonzero(t[..., 4] > 0)
    assert sorted(set(rows)) == [37, 38, 39] and sorted(set(cols)) == [37, 38, 39]
    centres = {(round(float((t[r, c, 0] + c) / 76 * 608), 3), round(float((t[r, c, 1] + r) / 76 * 608), 3))
               for r, c in zip(rows, cols)}
    assert len(centres) == 1                              # the property that test asserts
    np.testing.assert_allclose(list(centres)[0], (cx, cy), atol=2e-3)
    assert np.all(t[rows, cols, 5] == 1.0) and np.all(t[rows, cols, 8] == 1.0)   # anchor 0, class 0


def test_tf_semantics_differ_from_numpy_where_survey_says():
    anchors = _anchors()
    # asymmetric centre (140.5, 230): exact centre (no floor), fractions swapped
    box = np.array([[[100, 200, 181, 260, 3]]], np.float32)
    y_tf = O.encode_targets_tf_compat(box, (S, S), anchors, C)
    y_np = O.encode_targets(box, (S, S), anchors, C)
    layer = int(np.argmax(_layer_of(y_tf)))
    assert _layer_of(y_tf)[layer] == 9 and _layer_of(y_np)[layer] == 9
    G = y_tf[layer].shape[1]
    gx, gy = np.float32(140.5) * np.float32(G / S), np.float32(230.0) * np.float32(G / S)
    col, row = int(gx), int(gy)
    centre_tf = y_tf[layer][0, row, col]
    assert centre_tf[0] == np.float32(gy - row)           # channel 0 <- frac(cy)   (:3337)
    assert centre_tf[1] == np.float32(gx - col)           # channel 1 <- frac(cx)   (:3338)
    centre_np = y_np[layer][0, row, col]
    assert centre_np[0] == np.float32(np.float64(140.0) * G / S - col)   # floored centre, own fraction
    # neighbour (row-1, col+1): ki = -1 (row), kj = +1 (col)
    nb = y_tf[layer][0, row - 1, col + 1]
    assert nb[0] == np.float32(-1.0) + np.float32(gy - row)
    assert nb[1] == np.float32(1.0) + np.float32(gx - col)


def test_tf_last_box_wins_and_no_occupancy_rule():
    anchors = _anchors()
    # two boxes of the same size one cell apart on the coarse grid: 6 shared cells
    b = np.array([[[100, 200, 180, 260, 1], [132, 200, 212, 260, 2], [0, 0, 0, 0, 0]]], np.float32)
    y = O.encode_targets_tf_compat(b, (S, S), anchors, C)
    layer = int(np.argmax(_layer_of(y)))
    t = y[layer][0]
    A = len(anchors[layer])
    assert _layer_of(y)[layer] == 12                       # 9 + 9 - 6 shared: nothing is skipped
    owners = t[..., 5 + A + 1] + 2 * t[..., 5 + A + 2]     # 1 -> first box, 2 -> second box
    assert int((owners == 2).sum()) == 9                   # the later box owns all of its cells
    assert int((owners == 1).sum()) == 3
    # the NumPy encoder keeps the first box's cells once the second has written three
    y_np = O.encode_targets(b, (S, S), anchors, C)
    t_np = y_np[layer][0]
    owners_np = t_np[..., 5 + A + 1] + 2 * t_np[..., 5 + A + 2]
    assert int((owners_np == 2).sum()) < 9


def test_tf_class_out_of_range_and_borders():
    anchors = _anchors()
    b = np.array([[[0, 0, 9, 9, 80], [599, 599, 608, 608, -1]]], np.float32)
    y = O.encode_targets_tf_compat(b, (S, S), anchors, C)            # no AssertionError
    with pytest.raises(AssertionError):
        O.encode_targets(b, (S, S), anchors, C)
    tot = 0
    for l, t in enumerate(y):
        pos = t[0][t[0, ..., 4] > 0]
        tot += len(pos)
        assert np.all(pos[:, 5 + len(anchors[l]):] == 0)   # one_hot(out of range) = zeros
    assert tot == 4 + 4                                    # both boxes sit in a corner cell


@pytest.mark.gpu
@pytest.mark.parametrize("seed,B,N,S_,C_,layout", [
    (0, 16, 100, 608, 80, "uniform"), (1, 8, 300, 320, 80, "mosaic"),
    (2, 8, 20, 416, 20, "uniform"), (3, 4, 100, 512, 1, "mosaic")])
def test_gpu_tf_compat_matches_oracle(seed, B, N, S_, C_, layout):
    import torch
    from multigriddet_b200 import engine
    anchors = _anchors()
    boxes = synth.synth_boxes(seed, B, N, S_, C_, layout=layout, corners="frac")
    boxes[0, 0, 4] = C_ + 3                                 # out-of-range ids are legal here
    boxes[B - 1, 1, 4] = -2
    ref = O.encode_targets_tf_compat(boxes, (S_, S_), anchors, C_)
    for src in (boxes, torch.from_numpy(boxes).cuda()):
        got = engine.encode_targets(src, (S_, S_), anchors, C_, semantics="tf_compat")
        for g, r in zip(got, ref):
            g = g.cpu().numpy() if hasattr(g, "cpu") else g
            assert np.array_equal(g[..., 4:], r[..., 4:])                   # masks / one-hots
            assert np.array_equal(g[..., :2], r[..., :2])                   # offsets: exact float32
            np.testing.assert_allclose(g[..., 2:4], r[..., 2:4], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_gpu_tf_dropin_entry_point():
    from multigriddet_b200.data import tf_preprocess_true_boxes, preprocess_true_boxes
    anchors = _anchors()
    boxes = synth.synth_boxes(5, 4, 50, S, C, corners="frac")
    grids = [(19, 19), (38, 38), (76, 76)]
    ref = O.encode_targets_tf_compat(boxes, (S, S), anchors, C, grids)
    got = tf_preprocess_true_boxes(boxes, (S, S), anchors, C, False, grids)
    assert all(np.array_equal(g[..., 4:], r[..., 4:]) and np.array_equal(g[..., :2], r[..., :2])
               for g, r in zip(got, ref))
    same = tf_preprocess_true_boxes(boxes, (S, S), anchors, C, False, grids, semantics="numpy")
    base = preprocess_true_boxes(boxes, (S, S), anchors, C, False, grids)
    assert all(np.array_equal(a, b) for a, b in zip(same, base))


@pytest.mark.gpu
def test_gpu_tf_compat_on_odd_geometries():
    """The same generator as tests/test_oracle_vs_reference.py::
    test_tf_encoder_restatement_on_odd_geometries, CUDA against the oracle."""
    import torch
    from fuzz_util import tf_encoder_odd_case
    from multigriddet_b200 import engine
    rng = np.random.default_rng(78)
    for _ in range(40):
        S_, C_, anchors, grids, boxes = tf_encoder_odd_case(rng)
        ref = O.encode_targets_tf_compat(boxes, (S_, S_), anchors, C_, grids)
        got = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S_, S_), anchors, C_, grids,
                                    semantics="tf_compat")
        for g, r in zip(got, ref):
            g = g.cpu().numpy()
            assert np.array_equal(g[..., 4:], r[..., 4:]) and np.array_equal(g[..., :2], r[..., :2])
            np.testing.assert_allclose(g[..., 2:4], r[..., 2:4], rtol=1e-5, atol=1e-6)
