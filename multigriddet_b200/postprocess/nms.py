"""Drop-in for the reference's ``multigriddet/postprocess/nms.py`` greedy NMS classes.

``apply_nms`` keeps the reference's signature and return container (three lists of
one array each, detections in descending score order, ``([], [], [])`` on empty
input); the greedy pass runs in ``libmgd.so`` (``mgd_nms``).  Equal scores are taken
in ascending input position (the reference's ``argsort`` leaves that unspecified).
"""
from __future__ import annotations

import numpy as np

from .. import engine


class NMS:
    """Abstract base (reference nms.py:12-39)."""

    def __init__(self, use_iol: bool = False):
        self.use_iol = use_iol          # stored and never used, like the reference

    def apply_nms(self, boxes, classes, scores, nms_threshold, confidence):
        raise NotImplementedError("Subclasses must implement apply_nms method")


class _GreedyNMS(NMS):
    _method = "standard"

    def apply_nms(self, boxes, classes, scores, nms_threshold, confidence):
        if len(boxes) == 0:
            return [], [], []
        boxes = np.asarray(boxes)
        classes = np.asarray(classes)
        scores = np.asarray(scores)
        keep = engine.nms(boxes, scores, None, nms_threshold, self._method, per_class=False)
        if len(keep) == 0:
            return [], [], []
        return [boxes[keep]], [classes[keep]], [scores[keep]]


class StandardNMS(_GreedyNMS):
    """IoU greedy NMS (reference nms.py:83-148)."""
    _method = "standard"


class DIoUNMS(_GreedyNMS):
    """DIoU greedy NMS (reference nms.py:151-231)."""
    _method = "diou"


class ClusterNMS(_GreedyNMS):
    """Identical to StandardNMS in the reference (nms.py:320-385)."""
    _method = "cluster"


class SoftNMS(NMS):
    """Gaussian SoftNMS (reference nms.py:234-317): survivors in input order with their
    decayed scores."""

    def __init__(self, sigma: float = 0.5, score_threshold: float = 0.001):
        super().__init__()
        self.sigma = sigma
        self.score_threshold = score_threshold

    def apply_nms(self, boxes, classes, scores, nms_threshold, confidence):
        if len(boxes) == 0:
            return [], [], []
        boxes = np.asarray(boxes)
        classes = np.asarray(classes)
        keep, soft = engine.soft_nms(boxes, np.asarray(scores), self.sigma, self.score_threshold)
        if len(keep) == 0:
            return [], [], []
        return [boxes[keep]], [classes[keep]], [soft]


def nms_boxes(boxes, classes, scores, nms_threshold, use_iol=True, use_diou=False,
              confidence=0.5, is_soft=False, use_exp=False):
    """Backward-compatibility dispatcher (reference nms.py:389-399)."""
    if is_soft:
        nms = SoftNMS()
    elif use_diou:
        nms = DIoUNMS(use_iol=use_iol)
    else:
        nms = StandardNMS(use_iol=use_iol)
    return nms.apply_nms(boxes, classes, scores, nms_threshold, confidence)


def fast_cluster_nms_boxes(boxes, classes, scores, nms_threshold, use_iol=True, confidence=0.5):
    """Reference nms.py:402-405."""
    return ClusterNMS(use_iol=use_iol).apply_nms(boxes, classes, scores, nms_threshold, confidence)
