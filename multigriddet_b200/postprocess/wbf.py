"""Drop-in for the reference's ``multigriddet/postprocess/wbf.py``.

``WeightedBoxesFusion`` keeps the constructor, ``fuse_boxes`` signature and return
containers of the reference; clustering and fusion run in ``libmgd.so`` (``mgd_wbf``).
Equal scores within a class are visited in input order (the reference's ``argsort``
leaves that unspecified).
"""
from __future__ import annotations

import numpy as np

from .. import engine


class WeightedBoxesFusion:
    def __init__(self, iou_thr: float = 0.55, skip_box_thr: float = 0.0, conf_type: str = "avg",
                 allows_overflow: bool = False):
        self.iou_thr = iou_thr
        self.skip_box_thr = skip_box_thr
        self.conf_type = conf_type
        self.allows_overflow = allows_overflow          # stored and unused, like the reference

    def fuse_boxes(self, boxes_list, classes_list, scores_list, image_shape, weights=None):
        """Fuse the boxes of several models (reference wbf.py:38-128).  Returns three
        one-element lists ``([boxes], [classes], [scores])`` or ``([], [], [])``."""
        if len(boxes_list) == 0:
            return [], [], []
        if weights is None:
            weights = [1.0] * len(boxes_list)
        b, c, s, w = [], [], [], []
        for m, (boxes, classes, scores) in enumerate(zip(boxes_list, classes_list, scores_list)):
            if len(boxes) == 0:
                continue
            b.append(np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
            c.append(np.asarray(classes).reshape(-1))
            s.append(np.asarray(scores, dtype=np.float64).reshape(-1))
            w.append(np.full(len(b[-1]), float(weights[m])))
        if not b:
            return [], [], []
        classes_all = np.concatenate(c)
        fb, fs, fc = engine.wbf(np.concatenate(b), np.concatenate(s), classes_all, np.concatenate(w),
                                self.iou_thr, self.skip_box_thr, self.conf_type)
        if len(fb) == 0:
            return [], [], []
        return [fb], [fc.astype(classes_all.dtype)], [fs]


def weighted_boxes_fusion(boxes_list, classes_list, scores_list, image_shape, weights=None,
                          iou_thr=0.55, skip_box_thr=0.0, conf_type="avg", allows_overflow=False):
    """Backward-compatibility function (reference wbf.py:278-290)."""
    return WeightedBoxesFusion(iou_thr, skip_box_thr, conf_type, allows_overflow).fuse_boxes(
        boxes_list, classes_list, scores_list, image_shape, weights)
