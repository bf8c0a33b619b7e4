"""CPU restatement of the loss-side ignore mask of the reference
(``multigriddet/losses/multigrid_loss.py:445-492, 494-703``) -- SURVEY.md section 8(f)-2.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PINNED AGAINST THE REFERENCE'S OWN CODE OVER A TF-OP STAND-IN: TensorFlow is not installed in
this image, so the two reference methods are executed from their own source with
``oracle/tf_shim.py`` answering the ``tf.*`` / ``K.*`` calls in NumPy
(``ref_loader.load_tf_ignore_mask``); this restatement equals them bit for bit
(tests/test_ignore_mask.py; fixtures tests/golden/ignoremask_cases.npz).  The
transcendentals (``tanh``, ``sigmoid``, ``exp``) are NumPy's / libm's on both sides, which
TensorFlow does not promise to match bit for bit, so comparisons of the CUDA path use a 1e-5
tolerance and skip cells whose maximum IoU lies within 1e-5 of the ignore threshold.

Quirks of the reference that are restated as they are:
* ``tf.meshgrid(grid_x, grid_y, indexing='ij')`` (:547) makes the grid offset of tensor
  position [row i, col j] equal to (x = i, y = j): the ROW index is added to the x channel;
* anchors are multiplied by ``scale`` = input / grid as well (:570, :599), so box sizes are
  ``exp(t) * anchor * stride`` pixels;
* ground truth is the list of POSITIVE CELLS of the same layer (every object appears up to
  nine times), not the list of objects (:612-626).
"""
from __future__ import annotations

import numpy as np
from scipy.special import expit

_F = np.float32


def iou_center(boxes1, boxes2, eps=_F(1e-7)):
    """``_compute_iou_batch`` (:445-492) for boxes1 (P, 4), boxes2 (G, 4) in centre format,
    float32 -> (P, G)."""
    b1, b2 = boxes1[:, None, :], boxes2[None, :, :]
    mins1, maxs1 = b1[..., 0:2] - b1[..., 2:4] / _F(2.0), b1[..., 0:2] + b1[..., 2:4] / _F(2.0)
    mins2, maxs2 = b2[..., 0:2] - b2[..., 2:4] / _F(2.0), b2[..., 0:2] + b2[..., 2:4] / _F(2.0)
    wh = np.maximum(np.minimum(maxs1, maxs2) - np.maximum(mins1, mins2), _F(0.0))
    inter = wh[..., 0] * wh[..., 1]
    union = b1[..., 2] * b1[..., 3] + b2[..., 2] * b2[..., 3] - inter
    return inter / (union + eps)


def ignore_mask_layer(y_pred, y_true, anchors, input_shape, ignore_thresh=0.5, eps=1e-7):
    """One layer.  y_pred, y_true (B, G, G, 5 + A + C) float32; anchors (A, 2).
    Returns (ignore_mask, assigned_anchor_iou, max_iou_map), each (B, G, G, 1) float32."""
    y_pred = np.asarray(y_pred, dtype=_F)
    y_true = np.asarray(y_true, dtype=_F)
    anchors = np.asarray(anchors, dtype=_F)
    B, gh, gw = y_pred.shape[:3]
    A = anchors.shape[0]
    ii, jj = np.meshgrid(np.arange(gw, dtype=_F), np.arange(gh, dtype=_F), indexing="ij")   # :547
    grid = np.stack([ii, jj], -1)[None]                                                     # (1, gw, gh, 2)
    scale = np.array([_F(input_shape[1]) / _F(gw), _F(input_shape[0]) / _F(gh)], dtype=_F)   # (w, h) :552-555
    obj = (y_true[..., 4:5] > 0.5).astype(_F)                                               # :305
    true_xy_abs = (y_true[..., 0:2] + grid) * scale                                         # :558
    idx = np.argmax(y_true[..., 5:5 + A], axis=-1)                                          # :562
    true_wh_abs = np.exp(y_true[..., 2:4]) * anchors[idx] * scale                           # :570
    act = np.tanh(_F(0.15) * y_pred[..., 0:2]) + expit(_F(0.15) * y_pred[..., 0:2])          # :582
    pred_xy_abs = (act + grid) * scale                                                      # :585
    pred_wh_all = np.exp(y_pred[..., 2:4])[..., None, :] * anchors[None, None, None] * scale  # :599
    ignore = np.zeros((B, gh, gw, 1), _F)
    assigned = np.zeros((B, gh, gw, 1), _F)
    max_map = np.zeros((B, gh, gw, 1), _F)
    for b in range(B):
        valid = obj[b, ..., 0].reshape(-1) > 0.5
        pred = np.concatenate([np.repeat(pred_xy_abs[b][:, :, None, :], A, axis=2), pred_wh_all[b]], -1)
        pred = pred.reshape(-1, 4)
        if valid.any():
            gt = np.concatenate([true_xy_abs[b].reshape(-1, 2), true_wh_abs[b].reshape(-1, 2)], -1)[valid]
            iou = iou_center(pred, gt, _F(eps)).max(-1)                                     # :629-632
        else:
            iou = np.zeros(pred.shape[0], _F)
        iou = iou.reshape(gh, gw, A)
        mx = iou.max(-1)
        ignore[b, ..., 0] = ((mx > _F(ignore_thresh)) & (obj[b, ..., 0] < 0.5)).astype(_F)     # :684-688
        assigned[b, ..., 0] = np.take_along_axis(iou, idx[b][..., None], -1)[..., 0] * obj[b, ..., 0]   # :692-693
        max_map[b, ..., 0] = mx
    return ignore, assigned, max_map


def ignore_masks(y_preds, y_trues, anchors, input_shape, ignore_thresh=0.5, eps=1e-7):
    """All layers: lists of (ignore_mask, assigned_anchor_iou, max_iou_map) per layer."""
    return [ignore_mask_layer(p, t, a, input_shape, ignore_thresh, eps)
            for p, t, a in zip(y_preds, y_trues, anchors)]
