"""Randomised parity sweep: many small random geometries / knobs, CUDA path vs the C oracle.

Fixed seeds (the sweep is deterministic); every case is tiny, the value is in the breadth:
input sizes that are not multiples of 32, 1-5 layers, 1-8 anchors per layer with odd channel
counts (generic decode path), zero / interleaved padding rows, boxes on the border and
outside, thresholds at the extremes, per-class and every NMS method, both NMS kernels.
"""
import os

import numpy as np
import pytest

from multigriddet_b200 import engine, synth

pytestmark = pytest.mark.gpu


from fuzz_util import random_boxes as _random_boxes, random_head as _random_head, tie_free as _tie_free  # noqa: E402


@pytest.mark.parametrize("seed", range(48))
def test_random_geometry_encode_and_decode(c_oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    S, C, anchors, grids = _random_head(rng)
    B, N = int(rng.integers(1, 7)), int(rng.integers(1, 40))
    boxes = _tie_free(_random_boxes(rng, B, N, S, C), anchors)
    got, st = engine.encode_targets(boxes, (S, S), anchors, C, grids, return_stats=True)
    ref, rst = c_oracle.encode_targets(boxes, (S, S), anchors, C, grids, return_stats=True)
    assert st["n_valid_boxes"] == rst["n_valid_boxes"] and st["n_skipped_writes"] == rst["n_skipped_writes"]
    for g, r in zip(got, ref):
        assert np.array_equal(g[..., 4:], r[..., 4:]) and np.array_equal(g[..., :2], r[..., :2])
        np.testing.assert_allclose(g[..., 2:4], r[..., 2:4], rtol=1e-5, atol=1e-6)
    # head outputs: the targets planted with noise, plus a random dense component
    import torch
    if _planted_ok(anchors):
        preds = [p.numpy().copy() for p in synth.planted_head_outputs([torch.from_numpy(y) for y in ref],
                                                                      len(anchors[0]), seed)]
    else:
        preds = [rng.normal(0, 2, y.shape).astype(np.float32) for y in ref]
    for p in preds:
        p += rng.normal(0, float(rng.choice([0.0, 0.5, 2.0])), p.shape).astype(np.float32)
    shapes = np.stack([rng.integers(S // 2, 3 * S, B), rng.integers(S // 2, 3 * S, B)], 1).astype(np.int32)
    kw = dict(max_boxes=int(rng.choice([1, 5, 100, 300])), confidence=float(rng.choice([0.0, 0.001, 0.05, 0.3, 0.9])),
              nms_threshold=float(rng.choice([0.0, 0.3, 0.45, 0.7, 1.0])),
              nms_method=str(rng.choice(["diou", "standard", "cluster"])), per_class=bool(rng.integers(0, 2)),
              use_softmax=bool(rng.random() < 0.8), rescore_confidence=bool(rng.random() < 0.85))
    refd = c_oracle.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
    for force in ("1", "1000000"):
        os.environ["MGD_NMS_WARP_MIN_IMAGES"] = force
        try:
            gotd = engine.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
        finally:
            del os.environ["MGD_NMS_WARP_MIN_IMAGES"]
        assert np.array_equal(gotd["counts"], refd["counts"]), (seed, kw, force)
        for b in range(B):
            k = int(refd["counts"][b])
            assert np.array_equal(gotd["index"][b, :k], refd["index"][b, :k]), (seed, b, kw, force)
            assert np.array_equal(gotd["classes"][b, :k], refd["classes"][b, :k])
            assert np.array_equal(gotd["scores"][b, :k], refd["scores"][b, :k])
            np.testing.assert_allclose(gotd["boxes_xywh"][b, :k], refd["boxes_xywh"][b, :k], rtol=1e-5, atol=1e-4)


def _planted_ok(anchors):
    return len({len(a) for a in anchors}) == 1


@pytest.mark.parametrize("seed", range(12))
def test_explicit_nms_coordinate_regimes(seed):
    """mgd_nms on caller-supplied boxes: the float32 level of the pair test must never decide
    differently from the float64 formula, whatever the coordinate regime -- normalised [0, 1]
    boxes (where the 1e-8 terms matter), pixel boxes, boxes far from the origin (float32 cannot
    resolve them: everything must fall through to the exact test), identical and zero-area
    boxes, pairs engineered to sit right at the threshold."""
    from oracle import mgd_oracle as O
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(2, 700))
    scale, offset = [(1.0, 0.0), (600.0, 0.0), (600.0, 1.0e6), (1e-3, 0.0), (4000.0, -2000.0), (50.0, 3.0e4)][seed % 6]
    xy = rng.uniform(0, 1, (n, 2))
    wh = np.exp(rng.normal(np.log(0.12), 0.7, (n, 2))).clip(1e-4, 0.9)
    boxes = np.concatenate([xy, wh], 1)
    dup = rng.integers(0, n, n // 6)
    boxes[rng.integers(0, n, len(dup))] = boxes[dup]                       # identical boxes (IoU = 1)
    boxes[rng.integers(0, n, max(n // 20, 1)), 2:] = 0.0                    # zero-area boxes
    # pairs at the threshold: a shifted copy whose IoU is within ~1e-7 of thr
    thr = float(rng.choice([0.3, 0.45, 0.5, 0.7]))
    for _ in range(max(n // 10, 1)):
        i, j = rng.integers(0, n, 2)
        w = boxes[i, 2]
        if w <= 0:
            continue
        d = w * (1 - thr) / (1 + thr)                                       # IoU(d) = (w - d) / (w + d) = thr
        boxes[j] = boxes[i] + np.array([d, 0, 0, 0])
    boxes = boxes * scale
    boxes[:, :2] += offset
    scores = rng.uniform(0, 1, n)
    scores[rng.integers(0, n, n // 8)] = scores[rng.integers(0, n, n // 8)]    # ties
    classes = rng.integers(0, 3, n)
    for method, diou in (("diou", True), ("standard", False)):
        for per_class in (False, True):
            ref = O.greedy_nms(boxes, scores, thr, diou, classes=classes, per_class=per_class)
            got = engine.nms(boxes, scores, classes, thr, method, per_class)
            assert np.array_equal(got, ref), (seed, method, per_class, scale, offset)


@pytest.mark.parametrize("force", ["1", "1000000"])
def test_decode_extreme_image_shapes(c_oracle, force):
    """Letterbox targets from thumbnails to 48-megapixel frames: box coordinates span five
    orders of magnitude across the batch; both NMS kernels must still agree with the oracle."""
    import torch
    S, C, B = 416, 20, 8
    anchors = synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(77, B, 40, S, C)
    y = c_oracle.encode_targets(boxes, (S, S), anchors, C)
    preds = [p.numpy() for p in synth.planted_head_outputs([torch.from_numpy(t) for t in y], 3, 77)]
    shapes = np.array([[31, 47], [6000, 8000], [416, 416], [1, 1], [12000, 90], [480, 640], [8000, 6000], [75, 3000]], np.int32)
    kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")
    ref = c_oracle.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
    os.environ["MGD_NMS_WARP_MIN_IMAGES"] = force
    try:
        got = engine.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
    finally:
        del os.environ["MGD_NMS_WARP_MIN_IMAGES"]
    assert np.array_equal(got["counts"], ref["counts"])
    for b in range(B):
        k = int(ref["counts"][b])
        assert np.array_equal(got["index"][b, :k], ref["index"][b, :k]), (b, shapes[b])
        np.testing.assert_allclose(got["boxes_xywh"][b, :k], ref["boxes_xywh"][b, :k], rtol=1e-5, atol=1e-4)


def test_decode_non_finite_logits_are_dropped_like_the_reference(c_oracle):
    """NaN / +-inf in the objectness, anchor or class logits of a cell: the reference's
    softmax turns the score into NaN (or 0) and `score >= confidence` drops the cell."""
    import torch
    S, C, B = 416, 20, 4
    anchors = synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(78, B, 30, S, C)
    y = c_oracle.encode_targets(boxes, (S, S), anchors, C)
    preds = [p.numpy().copy() for p in synth.planted_head_outputs([torch.from_numpy(t) for t in y], 3, 78)]
    rng = np.random.default_rng(3)
    poisoned = 0
    for l, p in enumerate(preds):
        pos = np.argwhere(y[l][..., 4] > 0)
        for n, (b, i, j) in enumerate(pos[rng.permutation(len(pos))[:12]]):
            ch = [4, 5 + int(rng.integers(0, 3)), 8 + int(rng.integers(0, C))][n % 3]
            p[b, i, j, ch] = [np.nan, np.inf, -np.inf][(n // 3) % 3]
            poisoned += 1
    assert poisoned >= 24
    kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")
    with np.errstate(all="ignore"):
        ref = c_oracle.decode_nms(preds, [(S, S)], (S, S), anchors, C, **kw)
    got = engine.decode_nms(preds, (S, S), (S, S), anchors, C, **kw)
    assert np.array_equal(got["counts"], ref["counts"])
    for b in range(B):
        k = int(ref["counts"][b])
        assert np.array_equal(got["index"][b, :k], ref["index"][b, :k]), b
        assert np.array_equal(got["scores"][b, :k], ref["scores"][b, :k])
