"""Drop-in for ``MultiGridLoss._compute_ignore_mask`` (reference
``multigriddet/losses/multigrid_loss.py:494-703``), computed by ``mgd_ignore_mask``.

The three results are forward-only in the reference (a mask cast from a boolean, two
``stop_gradient`` IoU maps), so inside the TF loss they can come from
``tf.numpy_function`` / DLPack without touching the gradient path.  TensorFlow is not
installed in this image; parity is pinned against the reference methods' own source executed
over a NumPy stand-in for their ``tf.*`` / ``K.*`` ops (``tests/test_ignore_mask.py``,
``tests/golden/ignoremask_cases.npz``): masks exactly, IoU maps to 1e-5.
"""
from __future__ import annotations

import numpy as np

from .. import engine


def compute_ignore_masks(y_preds, y_trues, anchors, input_shape, num_classes, ignore_thresh=0.5,
                         eps=1e-7):
    """All layers at once: list of ``(ignore_mask, assigned_anchor_iou, max_iou_map)``, each
    ``(B, G, G, 1)`` float32, in the memory space of the inputs."""
    return engine.ignore_masks(y_preds, y_trues, anchors, input_shape, num_classes, ignore_thresh, eps)


def compute_ignore_mask(y_pred_layer, y_true_layer, anchors_layer, input_shape, ignore_thresh=0.5,
                        eps=1e-7):
    """One layer, the granularity the reference calls it at (:316): ``y_pred_layer`` /
    ``y_true_layer`` ``(B, G, G, 5 + A + C)``, ``anchors_layer`` ``(A, 2)`` pixels."""
    A = len(anchors_layer)
    C = int(y_pred_layer.shape[-1]) - 5 - A
    return compute_ignore_masks([y_pred_layer], [y_true_layer], [np.asarray(anchors_layer)],
                                input_shape, C, ignore_thresh, eps)[0]


def encode_and_ignore_masks(true_boxes, y_preds, anchors, input_shape, num_classes, ignore_thresh=0.5,
                            eps=1e-7, want_y_true=True):
    """``preprocess_true_boxes`` + ``_compute_ignore_mask`` of all layers in one library call on
    device tensors: the mask is fed by the encoder's owner table, the dense ``y_true`` is never
    re-read (and not even written with ``want_y_true=False``).  Returns ``(y_true, masks)``."""
    return engine.encode_ignore_masks(true_boxes, y_preds, anchors, input_shape, num_classes,
                                      ignore_thresh, eps, want_y_true)
