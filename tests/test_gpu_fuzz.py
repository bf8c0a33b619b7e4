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


def _random_head(rng):
    L = int(rng.integers(1, 6))
    S = int(rng.choice([64, 96, 100, 128, 160, 224, 250]))
    strides = (32, 16, 8, 4, 2)[:L]
    if S // strides[0] < 1:
        S = 64
    A = [int(rng.integers(1, 9)) if rng.random() < 0.3 else 3 for _ in range(L)]
    C = int(rng.choice([1, 2, 3, 7, 20, 80]))
    dt = np.float64 if rng.random() < 0.3 else np.float32
    anchors = [np.sort(rng.uniform(4, S * 0.9, (a, 2)), axis=0).astype(dt) for a in A]
    grids = [(max(S // s, 1), max(S // s, 1)) for s in strides]
    return S, C, anchors, grids


def _random_boxes(rng, B, N, S, C):
    boxes = np.zeros((B, N, 5), np.float32)
    for b in range(B):
        n = int(rng.integers(0, N + 1))
        rows = rng.permutation(N)[:n] if rng.random() < 0.5 else np.arange(n)
        c = rng.uniform(-0.1 * S, 1.1 * S, (n, 2))
        wh = np.exp(rng.normal(np.log(S / 6), 1.0, (n, 2))).clip(0.5, 1.5 * S)
        x1y1 = c - wh / 2
        x2y2 = c + wh / 2
        if rng.random() < 0.5:
            x1y1, x2y2 = np.floor(x1y1), np.ceil(x2y2)
        boxes[b, rows, :4] = np.concatenate([x1y1, x2y2], 1)
        boxes[b, rows, 4] = rng.integers(0, C, n)
    return boxes


def _tie_free(boxes, anchors):
    """Drop boxes whose two best rounded IoLs tie (the reference's argsort is host-dependent
    there): zero their rows."""
    table = np.concatenate(anchors, 0).astype(np.float64)
    wh = (boxes[..., 2:4] - boxes[..., 0:2]).astype(np.float64)
    inter = np.minimum(wh[..., None, 0], table[:, 0]) * np.minimum(wh[..., None, 1], table[:, 1])
    big = np.maximum((wh[..., 0] * wh[..., 1])[..., None], table[:, 0] * table[:, 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        iol = np.round(inter / big, 3)
    top2 = np.sort(iol, -1)[..., -2:] if table.shape[0] > 1 else None
    if top2 is not None:
        tie = np.abs(top2[..., 1] - top2[..., 0]) < 2e-3
        boxes[tie] = 0
    return boxes


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
