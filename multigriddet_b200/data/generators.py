"""Drop-in for the target-encoding entry points of the reference's
``multigriddet/data/generators.py``.

Same names, argument meaning, return containers and error behaviour as the
reference; the work happens in ``libmgd.so`` (``mgd_encode_targets``).
"""
from __future__ import annotations

import numpy as np

from .. import engine


def get_anchor_mask(anchors):
    """Global anchor indices grouped per layer (reference generators.py:2473-2483)."""
    mask, start = [], 0
    for layer in anchors:
        mask.append(list(range(start, start + len(layer))))
        start += len(layer)
    return mask


def _as_host_anchors(anchors):
    out = []
    for a in anchors:
        if hasattr(a, "numpy") and not isinstance(a, np.ndarray):
            a = a.numpy()                       # tf.constant / torch tensor
        out.append(np.asarray(a))
    return out


def preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes, multi_anchor_assign,
                          grid_shapes=None, iou_thresh=0.2):
    """Multi-grid ``y_true`` targets (reference generators.py:3393-3473).

    true_boxes: (B, N, 5) ``[x1, y1, x2, y2, class]`` in pixels, zero rows = padding.
    Returns a list of ``float32`` arrays ``(B, Gh, Gw, 5 + A + C)``, one per layer.
    ``AssertionError`` if a class id is >= ``num_classes`` (generators.py:3409).
    ``multi_anchor_assign`` and ``iou_thresh`` are accepted and ignored exactly as in
    the reference (it always calls ``best_fit_and_layer(..., False)``, :3435).

    NumPy in -> NumPy out (the library stages through the GPU); torch CUDA tensor
    in -> torch CUDA tensors out (zero-copy).
    """
    del multi_anchor_assign, iou_thresh
    anchors = _as_host_anchors(anchors)
    if grid_shapes is not None:
        grid_shapes = [(int(g[0]), int(g[1])) for g in grid_shapes]
    input_shape = (int(input_shape[0]), int(input_shape[1]))
    return engine.encode_targets(true_boxes, input_shape, anchors, int(num_classes), grid_shapes)


def tf_preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes, multi_anchor_assign,
                             grid_shapes, debug_aug_pipeline=False, semantics="tf_compat"):
    """Drop-in for the reference's ``@tf.function`` encoder (generators.py:2696-3390),
    its default training path.

    ``semantics="tf_compat"`` (default) reproduces what that function computes, which is
    NOT what ``preprocess_true_boxes`` computes: exact box centres,
    unrounded IoL with the Keras epsilon, every in-bounds cell of the 3x3 block written,
    the highest box index wins a contested cell (CPU ``tensor_scatter_nd_update`` order),
    xy stored as ``[-dcol + frac(cy), -drow + frac(cx)]``, no class-range error.  TensorFlow
    is not installed in this image; the semantics are pinned against the reference function's
    own source executed with a NumPy stand-in answering its ``tf.*`` calls (the test suite's
    ``oracle/tf_shim.py``; fixtures ``tests/golden/tfencode_*.npz``), logarithms to 1e-5.
    ``semantics="numpy"`` routes to the NumPy encoder's self-consistent rules instead.

    A ctypes library cannot be traced into a TF graph: inside ``dataset.map`` call
    it through ``tf.py_function`` / ``tf.numpy_function``.  TensorFlow tensors are
    accepted eagerly via DLPack when TF is present; NumPy and torch CUDA tensors
    always work.  Returns the same container type it was given.
    """
    del debug_aug_pipeline, multi_anchor_assign
    anchors = _as_host_anchors(anchors)
    anchors = [np.asarray(a, dtype=np.float32) for a in anchors]      # tf.constant(..., float32), :1446
    shape = tuple(int(v) for v in np.asarray(input_shape).reshape(-1)[:2])
    if grid_shapes is not None:
        grid_shapes = [(int(g[0]), int(g[1])) for g in grid_shapes]
    tf_mod = type(true_boxes).__module__.split(".")[0] == "tensorflow"
    if tf_mod:
        import tensorflow as tf            # only reachable where TF exists
        boxes = np.from_dlpack(tf.experimental.dlpack.to_dlpack(true_boxes)) \
            if hasattr(np, "from_dlpack") else true_boxes.numpy()
        y = engine.encode_targets(boxes, shape, anchors, int(num_classes), grid_shapes,
                                  semantics=semantics)
        return [tf.convert_to_tensor(t) for t in y]
    return engine.encode_targets(true_boxes, shape, anchors, int(num_classes), grid_shapes,
                                 semantics=semantics)


def expand_box_capacity(boxes_dense, mosaic_enabled=False, mixup_enabled=False):
    """``_expand_box_capacity`` of the reference's tf.data path (generators.py:1983-2034): pad the
    (B, N, 5) box tensor with zero rows to N x {1, 2, 4, 8} for none / MixUp / Mosaic / both.
    Shape glue (no arithmetic): NumPy in -> NumPy out, torch in -> torch out."""
    factor = 8 if (mosaic_enabled and mixup_enabled) else 4 if mosaic_enabled else 2 if mixup_enabled else 1
    if factor == 1:
        return boxes_dense
    B, N = int(boxes_dense.shape[0]), int(boxes_dense.shape[1])
    if engine._is_torch(boxes_dense):
        import torch
        pad = torch.zeros((B, N * (factor - 1), 5), dtype=boxes_dense.dtype, device=boxes_dense.device)
        return torch.cat([boxes_dense, pad], dim=1)
    return np.concatenate([np.asarray(boxes_dense),
                           np.zeros((B, N * (factor - 1), 5), dtype=np.asarray(boxes_dense).dtype)], axis=1)


def letterbox_boxes(boxes, src_shapes, input_shape, max_boxes_per_image, counts=None,
                    mosaic_enabled=False, mixup_enabled=False, multiscale_shapes=None, hflip=None):
    """Box side of ``build_tf_dataset`` up to the encoder's input, for a batch: the letterbox /
    multi-scale transform of ``_preprocess_image_and_boxes`` (generators.py:1859-1916), the
    flip of ``tf_random_horizontal_flip`` (:227-256, coin supplied by the caller), ``padded_batch``
    to ``max_boxes_per_image`` (:1963-1976) and ``_expand_box_capacity`` (:1983-2034), in one
    kernel (``mgd_letterbox_boxes``).  TensorFlow is not installed here; parity is pinned, bit
    for bit, against those reference functions' own source executed over a NumPy stand-in for
    their ``tf.*`` ops (``tests/test_box_transforms.py``, ``tests/golden/tfboxes_cases.npz``)."""
    factor = 8 if (mosaic_enabled and mixup_enabled) else 4 if mosaic_enabled else 2 if mixup_enabled else 1
    return engine.letterbox_boxes_batch(boxes, src_shapes, input_shape, max_boxes_per_image, counts,
                                        factor, multiscale_shapes, hflip)
