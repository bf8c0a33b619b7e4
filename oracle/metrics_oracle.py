"""CPU restatement of the reference's detection-to-ground-truth matching and AP
arithmetic (``multigriddet/evaluation/metrics.py``) -- SURVEY.md section 8(f)-4.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity status: PINNED against
the reference executed from ``/root/reference`` (``tests/test_oracle_vs_reference.py``)
and against ``tests/golden/metrics_cases.npz``.

Works on the tensor layout the CUDA path uses (padded per-image detections as
``mgd_decode_nms`` emits them, padded per-image ground truth) instead of the
reference's lists of dicts; ``to_dicts`` builds the reference's containers from it.

Deterministic choice where the reference is unspecified: ``np.argsort(scores)[::-1]``
(``metrics.py:93``) is unstable for equal scores; here ties are ordered like a
reversed *stable* argsort (later prediction first).
"""
from __future__ import annotations

import numpy as np


def iou_corner(p, g):
    """``calculate_iou_matrix`` (metrics.py:28-70) for one pair of xyxy boxes, float64."""
    x1 = max(p[0], g[0]); y1 = max(p[1], g[1])
    x2 = min(p[2], g[2]); y2 = min(p[3], g[3])
    inter = max(0.0, x2 - x1) * max(0.0, y2 - y1)
    union = (p[2] - p[0]) * (p[3] - p[1]) + (g[2] - g[0]) * (g[3] - g[1]) - inter
    return inter / union if union > 0 else 0.0


def iou_centre(p, g):
    """``BoxUtils.box_iou`` (utils/boxes.py:16-58) as the un-cached matcher calls it
    (metrics.py:128): the four numbers are read as ``[x, y, w, h]`` centre format even
    though the evaluator stores xyxy corners -- restated as is."""
    x1, y1, w1, h1 = p
    x2, y2, w2, h2 = g
    a_x0, a_y0, a_x1, a_y1 = x1 - w1 / 2, y1 - h1 / 2, x1 + w1 / 2, y1 + h1 / 2
    b_x0, b_y0, b_x1, b_y1 = x2 - w2 / 2, y2 - h2 / 2, x2 + w2 / 2, y2 + h2 / 2
    ix0, iy0 = max(a_x0, b_x0), max(a_y0, b_y0)
    ix1, iy1 = min(a_x1, b_x1), min(a_y1, b_y1)
    if ix1 <= ix0 or iy1 <= iy0:
        return 0.0
    inter = (ix1 - ix0) * (iy1 - iy0)
    union = w1 * h1 + w2 * h2 - inter
    return inter / union if union > 0 else 0.0


def match_image(det_boxes, det_scores, det_classes, gt_boxes, gt_classes, thresholds, cached=True):
    """TP flags of one image's detections for every IoU threshold.

    Restates the per-(image, class) effect of ``match_predictions_to_gt_cached``
    (metrics.py:147-218, ``cached=True``: corner IoU, a ground truth is a candidate only
    if its IoU beats 0.0, :196-199) or ``match_predictions_to_gt`` (:73-144,
    ``cached=False``: centre-format IoU, first maximum even at IoU 0, :133-135).
    Detections are visited in descending score (ties: later slot first); each claims the
    best still-unmatched ground truth of its class when that IoU >= threshold.
    Returns (tp (T, n) uint8 in slot order, matched_gt (T, n) int32, -1 = none).
    """
    n, g = len(det_scores), len(gt_classes)
    T = len(thresholds)
    tp = np.zeros((T, n), np.uint8)
    who = np.full((T, n), -1, np.int32)
    order = np.argsort(np.asarray(det_scores), kind="stable")[::-1]
    iou = iou_corner if cached else iou_centre
    for t, thr in enumerate(thresholds):
        taken = np.zeros(g, bool)
        for d in order:
            best, best_j = 0.0, -1
            first = True
            for j in range(g):
                if gt_classes[j] != det_classes[d] or taken[j]:
                    continue
                v = iou([float(x) for x in det_boxes[d]], [float(x) for x in gt_boxes[j]])
                if cached:
                    if v > best:
                        best, best_j = v, j
                else:
                    if first or v > best:
                        best, best_j = v, j
                    first = False
            if best_j >= 0 and best >= thr:
                tp[t, d] = 1
                who[t, d] = best_j
                taken[best_j] = True
    return tp, who


def match_batch(det_boxes, det_scores, det_classes, det_counts, gt_boxes, gt_classes, gt_counts,
                thresholds, cached=True):
    """``match_image`` over padded batches: det (B, M, ...), gt (B, N, ...)."""
    B, M = det_scores.shape
    T = len(thresholds)
    tp = np.zeros((T, B, M), np.uint8)
    who = np.full((T, B, M), -1, np.int32)
    for b in range(B):
        n, g = int(det_counts[b]), int(gt_counts[b])
        a, w = match_image(det_boxes[b, :n], det_scores[b, :n], det_classes[b, :n],
                           gt_boxes[b, :g], gt_classes[b, :g], thresholds, cached)
        tp[:, b, :n] = a
        who[:, b, :n] = w
    return tp, who


def precision_recall(tp_sorted, num_gt):
    """``compute_precision_recall`` (metrics.py:221-246); tp_sorted in descending-score order."""
    if len(tp_sorted) == 0:
        return np.array([0.0]), np.array([0.0])
    cum_tp = np.cumsum(tp_sorted.astype(bool))
    cum_fp = np.cumsum(~tp_sorted.astype(bool))
    return cum_tp / (cum_tp + cum_fp + 1e-8), cum_tp / (num_gt + 1e-8)


def average_precision(precisions, recalls, method="coco"):
    """``compute_average_precision`` (metrics.py:249-300)."""
    if len(precisions) == 0 or len(recalls) == 0:
        return 0.0
    if method == "voc":
        out = []
        for r in np.arange(0, 1.1, 0.1):
            sel = precisions[recalls >= r]
            out.append(np.max(sel) if len(sel) else 0.0)
        return np.mean(out)
    if method != "coco":
        raise ValueError(f"Unknown method: {method}")
    idx = np.argsort(recalls)
    r, p = recalls[idx], precisions[idx]
    interp = np.maximum.accumulate(p[::-1])[::-1]
    if len(r) > 1:
        return float(np.sum((r[1:] - r[:-1]) * (interp[1:] + interp[:-1]) / 2.0))   # np.trapz
    return float(interp[0] * r[0])


def class_ap(tp_flat, scores_flat, classes_flat, gt_classes_flat, class_id, method="coco"):
    """``calculate_ap_for_class(_cached)`` (metrics.py:303-385) given the flags of ALL
    detections (flat, any order) for one threshold."""
    sel = classes_flat == class_id
    n_gt = int(np.sum(gt_classes_flat == class_id))
    if not sel.any():
        return 0.0 if n_gt > 0 else 1.0
    if n_gt == 0:
        return 0.0
    order = np.argsort(scores_flat[sel], kind="stable")[::-1]
    p, r = precision_recall(tp_flat[sel][order], n_gt)
    return average_precision(p, r, method)


def to_dicts(det_boxes, det_scores, det_classes, det_counts, gt_boxes, gt_classes, gt_counts):
    """The reference's containers (lists of dicts, evaluator.py:290-297, 395-399)."""
    preds, gts = [], []
    for b in range(len(det_counts)):
        for i in range(int(det_counts[b])):
            preds.append({"image_id": b, "class": int(det_classes[b, i]), "score": float(det_scores[b, i]),
                          "bbox": [v.item() for v in det_boxes[b, i]]})
        for j in range(int(gt_counts[b])):
            gts.append({"image_id": b, "class": int(gt_classes[b, j]),
                        "bbox": [float(v) for v in gt_boxes[b, j]]})
    return preds, gts
