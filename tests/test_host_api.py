"""Host-side logic of the drop-in modules (no GPU needed)."""
import numpy as np
import pytest

from multigriddet_b200 import sharding, synth
from multigriddet_b200.data import MultiGridConfig, MultiGridTargetEncoder, get_anchor_mask
from multigriddet_b200.postprocess import (DIoUNMS, MultiGridDecoder, NMS, SoftNMS, nms_boxes)
from oracle import mgd_oracle as O


def test_anchor_mask_matches_reference_layout():
    assert get_anchor_mask(synth.coco_anchors()) == [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
    assert get_anchor_mask([np.ones((2, 2)), np.ones((4, 2))]) == [[0, 1], [2, 3, 4, 5]]


def test_decoder_shim_conventions():
    anchors = synth.coco_anchors(np.float32)
    dec = MultiGridDecoder(anchors, 80)
    assert dec.num_layers == 3 and dec.use_softmax and dec.rescore_confidence
    with pytest.raises(ValueError, match="Expected 3 predictions, got 1"):
        dec.postprocess([np.zeros((1, 19, 19, 88), np.float32)], (608, 608), (608, 608))
    with pytest.raises(ValueError):
        dec.decode_predictions([np.zeros((1, 19, 19, 88), np.float32)] * 2)
    with pytest.raises(NotImplementedError):
        dec.postprocess([np.zeros((1, g, g, 88), np.float32) for g in (19, 38, 76)],
                        (608, 608), (608, 608), nms_method="wbf-something")
    # an empty scale: the reference returns three empty arrays
    b, c, s = dec.postprocess([np.zeros((0, g, g, 88), np.float32) for g in (19, 38, 76)],
                              (608, 608), (608, 608))
    assert b.size == c.size == s.size == 0
    # the kernel has ONE model size: the reference's two (input_shape for wh, model_image_size
    # for the letterbox) must agree, otherwise results would silently differ
    with pytest.raises(ValueError, match="input_shape"):
        dec.postprocess([np.zeros((1, g, g, 88), np.float32) for g in (13, 26, 52)],
                        (416, 416), (416, 416))
    # nothing above the threshold needs no device either
    b, c, s = dec.handle_predictions(np.zeros((1, 10, 85)), (608, 608), confidence=0.5)
    assert b.size == 0


def test_correct_boxes_matches_oracle():
    rng = np.random.default_rng(0)
    dec = MultiGridDecoder(synth.coco_anchors(np.float32), 20, input_shape=(416, 416))
    pred = rng.random((2, 50, 25))
    for shape in ((416, 416), (480, 640), (1080, 1920), (375, 500)):
        assert np.array_equal(dec.correct_boxes(pred, shape, (416, 416)),
                              O.correct_boxes(pred, shape, (416, 416)))
    xywh = rng.random((7, 4)) * 300
    assert np.array_equal(dec._convert_to_xyxy(xywh, (200, 250)), O.to_xyxy(xywh, (200, 250)))


def test_nms_shim_conventions():
    assert DIoUNMS().apply_nms(np.zeros((0, 4)), np.zeros(0), np.zeros(0), 0.5, 0.1) == ([], [], [])
    assert nms_boxes([], [], [], 0.5) == ([], [], [])
    with pytest.raises(NotImplementedError):
        NMS().apply_nms(np.zeros((1, 4)), np.zeros(1), np.ones(1), 0.5, 0.1)
    assert SoftNMS().apply_nms(np.zeros((0, 4)), np.zeros(0), np.zeros(0), 0.5, 0.1) == ([], [], [])
    assert SoftNMS().sigma == 0.5 and SoftNMS().score_threshold == 0.001 and DIoUNMS(use_iol=True).use_iol is True


def test_target_encoder_shim_fields():
    cfg = MultiGridConfig()
    assert cfg.input_shape == (608, 608) and cfg.num_classes == 80 and cfg.max_boxes == 100
    enc = MultiGridTargetEncoder(cfg)
    assert enc.grid_shapes == [(19, 19), (38, 38), (76, 76)] and enc.num_layers == 3


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 64, 4096, 4099):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)


def test_synth_is_deterministic_and_tie_free():
    a = synth.synth_boxes(3, 4, 50, 608, 80, layout="mosaic", corners="frac")
    b = synth.synth_boxes(3, 4, 50, 608, 80, layout="mosaic", corners="frac")
    assert np.array_equal(a, b) and a.dtype == np.float32
    assert synth.count_iol_ties(a, synth.coco_anchors()) == 0
    valid = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1]) > 0
    assert valid.any() and (a[valid][:, 4] < 80).all() and (a[..., :4] >= 0).all() and (a[..., :4] <= 608).all()
    shapes = synth.image_shapes(1, 16)
    assert shapes.shape == (16, 2) and shapes.dtype == np.int32


def test_pinned_pool_hands_out_plain_arrays_without_a_gpu():
    from multigriddet_b200 import _lib
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    pool = _lib.PinnedPool(cap_bytes=1 << 30)
    a = pool.empty((512, 512), np.float32)           # mgd_host_alloc -> MGD_ERR_NO_DEVICE -> np.empty
    assert a.shape == (512, 512) and a.dtype == np.float32 and pool._total == 0
    a[:] = 1.0


def test_exchange_layout_is_aligned_and_disjoint():
    """sharding.exchange_layout: whole-batch detection tensors inside an exchange buffer."""
    from multigriddet_b200 import sharding
    for n_total, m in ((1, 1), (11, 50), (4096, 100), (7, 3)):
        layout, total = sharding.exchange_layout(n_total, m)
        spans = []
        for name, (off, shape, dt) in layout.items():
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            assert off % 256 == 0 and shape[0] == n_total
            spans.append((off, off + nbytes))
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total
        assert set(layout) == {"boxes_xywh", "boxes_xyxy", "scores", "classes", "index", "counts"}
        # the two box tensors stay 16-byte aligned for every shard start (rows are 32 / 16 bytes)
        for lo in range(n_total):
            assert (layout["boxes_xywh"][0] + lo * m * 32) % 16 == 0
            assert (layout["boxes_xyxy"][0] + lo * m * 16) % 16 == 0
