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
