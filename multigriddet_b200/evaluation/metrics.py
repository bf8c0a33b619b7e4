"""Drop-in for ``multigriddet/evaluation/metrics.py`` of the reference: same function
names, arguments, return containers and conventions; the IoU and the greedy
detection-to-ground-truth matching (the O(P x G) part, a Python double loop in the
reference) run in ``libmgd.so`` (``mgd_match_detections``, ``mgd_iou_matrix``).  The
precision / recall / AP arithmetic on the resulting flag vectors is a few NumPy
reductions per class, exactly as in the reference.

Tie rule (the reference's ``np.argsort(scores)[::-1]``, metrics.py:93, is unstable):
equal scores are ordered like a reversed stable argsort, i.e. later prediction first.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

from .. import engine

COCO_THRESHOLDS = [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95]


def calculate_iou_matrix(boxes1: np.ndarray, boxes2: np.ndarray) -> np.ndarray:
    """(N, M) IoU of xyxy boxes -- metrics.py:28-70."""
    boxes1, boxes2 = np.asarray(boxes1), np.asarray(boxes2)
    if len(boxes1) == 0 or len(boxes2) == 0:
        return np.zeros((len(boxes1), len(boxes2)))
    if boxes1.shape[1] != 4 or boxes2.shape[1] != 4:
        raise ValueError("Boxes must have 4 coordinates")
    return engine.iou_matrix(boxes1, boxes2)


class _Packed:
    """Lists of dicts -> the padded per-image tensors the library takes."""

    def __init__(self, predictions: List[Dict], ground_truths: List[Dict]):
        ids = {}
        for rec in list(predictions) + list(ground_truths):
            ids.setdefault(rec["image_id"], len(ids))
        B = max(len(ids), 1)
        p_img = np.array([ids[p["image_id"]] for p in predictions], dtype=np.int64)
        g_img = np.array([ids[g["image_id"]] for g in ground_truths], dtype=np.int64)
        self.det_counts = np.bincount(p_img, minlength=B).astype(np.int32)
        self.gt_counts = np.bincount(g_img, minlength=B).astype(np.int32)
        M = max(int(self.det_counts.max(initial=0)), 1)
        N = max(int(self.gt_counts.max(initial=0)), 1)

        def slots(img, counts):
            order = np.argsort(img, kind="stable")
            start = np.concatenate([[0], np.cumsum(counts)[:-1]])
            slot = np.empty(len(img), dtype=np.int64)
            slot[order] = np.arange(len(img)) - np.repeat(start, counts)
            return slot
        self.p_img, self.p_slot = p_img, slots(p_img, self.det_counts)
        g_slot = slots(g_img, self.gt_counts)
        self.det_boxes = np.zeros((B, M, 4)); self.det_scores = np.zeros((B, M))
        self.det_classes = np.full((B, M), -1, dtype=np.int32)
        self.gt_boxes = np.zeros((B, N, 4)); self.gt_classes = np.full((B, N), -2, dtype=np.int32)
        if len(predictions):
            self.det_boxes[p_img, self.p_slot] = np.array([p["bbox"] for p in predictions], dtype=np.float64)
            self.det_scores[p_img, self.p_slot] = np.array([p["score"] for p in predictions], dtype=np.float64)
            self.det_classes[p_img, self.p_slot] = np.array([p["class"] for p in predictions])
        if len(ground_truths):
            self.gt_boxes[g_img, g_slot] = np.array([g["bbox"] for g in ground_truths], dtype=np.float64)
            self.gt_classes[g_img, g_slot] = np.array([g["class"] for g in ground_truths])
        self.pred_scores = np.array([p["score"] for p in predictions], dtype=np.float64)
        self.pred_classes = np.array([p["class"] for p in predictions], dtype=np.int64)
        self.gt_class_flat = np.array([g["class"] for g in ground_truths], dtype=np.int64)

    def flags(self, thresholds, mode) -> np.ndarray:
        """(T, P) uint8 TP flags in the order of the ``predictions`` list."""
        tp = engine.match_detections(self.det_boxes, self.det_scores, self.det_classes, self.det_counts,
                                     self.gt_boxes, self.gt_classes, self.gt_counts, thresholds,
                                     iou_mode=mode)
        return tp[:, self.p_img, self.p_slot] if len(self.p_img) else np.zeros((len(thresholds), 0), np.uint8)


def _sorted_flags(tp: np.ndarray, scores: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    order = np.argsort(scores, kind="stable")[::-1]
    flags = tp[order].astype(bool)
    return flags, ~flags, scores[order]


def match_predictions_to_gt(predictions: List[Dict], ground_truths: List[Dict],
                            iou_threshold: float = 0.5) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(tp_flags, fp_flags, scores) in descending-score order -- metrics.py:73-144 (the
    un-cached matcher: ``BoxUtils.box_iou`` reads the boxes as centre format).  Like the
    reference it matches on 'class' and 'image_id' of whatever lists it is given."""
    if len(predictions) == 0:
        return np.array([]), np.array([]), np.array([])
    pk = _Packed(predictions, ground_truths)
    return _sorted_flags(pk.flags([iou_threshold], "centre")[0], pk.pred_scores)


def match_predictions_to_gt_cached(predictions: List[Dict], ground_truths: List[Dict],
                                   iou_threshold: float, iou_cache=None):
    """metrics.py:147-218.  ``iou_cache`` is accepted for signature compatibility and not
    read: the corner-format IoUs it holds are recomputed on the device."""
    if len(predictions) == 0:
        return np.array([]), np.array([]), np.array([])
    pk = _Packed(predictions, ground_truths)
    return _sorted_flags(pk.flags([iou_threshold], "corner")[0], pk.pred_scores)


def compute_iou_cache_for_class(predictions: List[Dict], ground_truths: List[Dict], class_id: int):
    """metrics.py:456-514: ``{(pred_idx, gt_idx): iou}`` for same-image pairs of one class
    (indices into the class-filtered lists)."""
    cp = [p for p in predictions if p["class"] == class_id]
    cg = [g for g in ground_truths if g["class"] == class_id]
    if not cp or not cg:
        return {}
    cache = {}
    by_img_p, by_img_g = {}, {}
    for i, p in enumerate(cp):
        by_img_p.setdefault(p["image_id"], []).append(i)
    for j, g in enumerate(cg):
        by_img_g.setdefault(g["image_id"], []).append(j)
    for img, pi in by_img_p.items():
        gi = by_img_g.get(img)
        if not gi:
            continue
        mat = calculate_iou_matrix(np.array([cp[i]["bbox"] for i in pi]), np.array([cg[j]["bbox"] for j in gi]))
        for a, i in enumerate(pi):
            for b, j in enumerate(gi):
                cache[(i, j)] = float(mat[a, b])
    return cache


def compute_precision_recall(tp_flags: np.ndarray, fp_flags: np.ndarray, num_gt: int):
    """metrics.py:221-246."""
    if len(tp_flags) == 0:
        return np.array([0.0]), np.array([0.0])
    cum_tp, cum_fp = np.cumsum(tp_flags), np.cumsum(fp_flags)
    return cum_tp / (cum_tp + cum_fp + 1e-8), cum_tp / (num_gt + 1e-8)


def compute_average_precision(precisions: np.ndarray, recalls: np.ndarray, method: str = "coco") -> float:
    """metrics.py:249-300 ('coco': all-point interpolation + trapezoid; 'voc': 11 points)."""
    if len(precisions) == 0 or len(recalls) == 0:
        return 0.0
    if method == "voc":
        vals = []
        for r in np.arange(0, 1.1, 0.1):
            sel = precisions[recalls >= r]
            vals.append(np.max(sel) if len(sel) > 0 else 0.0)
        return np.mean(vals)
    if method == "coco":
        idx = np.argsort(recalls)
        r, p = recalls[idx], precisions[idx]
        interp = np.maximum.accumulate(p[::-1])[::-1]
        if len(r) > 1:
            return float(np.sum((r[1:] - r[:-1]) * (interp[1:] + interp[:-1]) / 2.0))
        return interp[0] * r[0]
    raise ValueError(f"Unknown method: {method}")


def _class_ap(tp_row, pk: _Packed, class_id: int, method: str) -> float:
    sel = pk.pred_classes == class_id
    n_gt = int(np.sum(pk.gt_class_flat == class_id))
    if not sel.any():
        return 0.0 if n_gt > 0 else 1.0
    if n_gt == 0:
        return 0.0
    tp, fp, _ = _sorted_flags(tp_row[sel], pk.pred_scores[sel])
    return compute_average_precision(*compute_precision_recall(tp, fp, n_gt), method)


def calculate_ap_for_class(predictions, ground_truths, class_id: int, iou_threshold: float = 0.5,
                           method: str = "coco") -> float:
    """metrics.py:303-341 (un-cached matcher)."""
    pk = _Packed(predictions, ground_truths)
    return _class_ap(pk.flags([iou_threshold], "centre")[0], pk, class_id, method)


def calculate_ap_for_class_cached(predictions, ground_truths, class_id: int, iou_threshold: float,
                                  iou_cache=None, method: str = "coco") -> float:
    """metrics.py:344-385 (cached matcher; ``iou_cache`` not read)."""
    pk = _Packed(predictions, ground_truths)
    return _class_ap(pk.flags([iou_threshold], "corner")[0], pk, class_id, method)


def _box_area(bbox) -> float:
    x1, y1, x2, y2 = bbox
    return (x2 - x1) * (y2 - y1)


def _filter_by_area(predictions, ground_truths, min_area=None, max_area=None):
    keep = lambda r: ((min_area is None or _box_area(r["bbox"]) >= min_area) and
                      (max_area is None or _box_area(r["bbox"]) < max_area))
    return [p for p in predictions if keep(p)], [g for g in ground_truths if keep(g)]


def calculate_map(predictions: List[Dict], ground_truths: List[Dict], num_classes: int,
                  iou_thresholds: List[float] = None, class_names: List[str] = None,
                  method: str = "coco", use_parallel: bool = True, optimize_classes: bool = True,
                  cache_ious: bool = True, compute_per_scale: bool = True) -> Dict[str, Any]:
    """mAP over classes and IoU thresholds -- metrics.py:541-815, same result keys.

    One ``mgd_match_detections`` launch covers every class and threshold.  Which matcher
    (and with it which IoU formula) the reference would have used is reproduced: the cached
    corner-IoU matcher by default; the un-cached centre-format one when ``cache_ious`` is
    false, when the parallel branch sees more than 10000 predictions (:604-606), and for
    the per-scale APs (:745-800).  ``use_parallel`` only takes part in that decision.
    """
    if iou_thresholds is None:
        iou_thresholds = list(COCO_THRESHOLDS)
    if class_names is None:
        class_names = [f"class_{i}" for i in range(num_classes)]
    results = {"mAP": 0.0, "mAP50": 0.0, "mAP75": 0.0, "per_class": {}, "per_iou": {},
               "num_predictions": len(predictions), "num_ground_truths": len(ground_truths)}
    if optimize_classes:
        active = sorted(set(p["class"] for p in predictions) | set(g["class"] for g in ground_truths))
    else:
        active = list(range(num_classes))
    if use_parallel and len(active) > 1 and cache_ious and len(predictions) > 10000:
        cache_ious = False
    pk = _Packed(predictions, ground_truths)
    tp = pk.flags(iou_thresholds, "corner" if cache_ious else "centre") if iou_thresholds else None
    iou_aps = {thr: [] for thr in iou_thresholds}
    class_aps = {}
    for class_id in active:
        name = class_names[class_id] if class_id < len(class_names) else f"class_{class_id}"
        res = {}
        for t, thr in enumerate(iou_thresholds):
            ap = _class_ap(tp[t], pk, class_id, method)
            res[f"AP{thr:.2f}"] = ap
            iou_aps[thr].append(ap)
        res["AP"] = np.mean(list(res.values()))
        class_aps[name] = res
    results["per_class"] = class_aps
    for thr in iou_thresholds:
        if len(iou_aps[thr]) > 0:
            results["per_iou"][f"mAP{thr:.2f}"] = np.mean(iou_aps[thr])
    if 0.5 in iou_thresholds:
        results["mAP50"] = results["per_iou"].get("mAP0.50", 0.0)
    if 0.75 in iou_thresholds:
        results["mAP75"] = results["per_iou"].get("mAP0.75", 0.0)
    if len(iou_thresholds) > 0:
        results["mAP"] = np.mean([results["per_iou"].get(f"mAP{thr:.2f}", 0.0) for thr in iou_thresholds])
    for key, lo, hi in (("APS", None, 1024.0), ("APM", 1024.0, 9216.0), ("APL", 9216.0, None)):
        results[key], results[key + "50"] = 0.0, 0.0
        if compute_per_scale:
            sp, sg = _filter_by_area(predictions, ground_truths, lo, hi)
            if len(sg) > 0:
                sub = calculate_map(sp, sg, num_classes, iou_thresholds, class_names, method,
                                    use_parallel=False, optimize_classes=optimize_classes,
                                    cache_ious=False, compute_per_scale=False)
                results[key], results[key + "50"] = sub["mAP"], sub.get("mAP50", 0.0)
    return results
