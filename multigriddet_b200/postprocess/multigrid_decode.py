"""Drop-in for the reference's ``multigriddet/postprocess/multigrid_decode.py``.

``MultiGridDecoder`` keeps the constructor, method names, argument meaning, return
containers and error behaviour of the reference class; ``postprocess`` is one call
into ``libmgd.so`` (``mgd_decode_nms``: decode -> letterbox -> threshold -> NMS ->
top-k -> xyxy, all on the GPU).  Safe to call from several threads on one decoder
object (the reference evaluator uses up to 8, evaluator.py:283-286).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .. import engine
from .nms import NMS, ClusterNMS, DIoUNMS, SoftNMS, StandardNMS
from .wbf import WeightedBoxesFusion

_EMPTY = lambda: (np.array([]), np.array([]), np.array([]))     # multigrid_decode.py:273-274


class MultiGridDecoder:
    def __init__(self, anchors: List[np.ndarray], num_classes: int,
                 input_shape: Tuple[int, int] = (608, 608), rescore_confidence: bool = True,
                 use_softmax: bool = True):
        self.anchors = anchors
        self.num_classes = num_classes
        self.input_shape = input_shape
        self.rescore_confidence = rescore_confidence
        self.use_softmax = use_softmax
        self.num_layers = len(anchors)

    # ---- dense API ---------------------------------------------------------------
    def _usable(self, predictions):
        if len(predictions) != self.num_layers:                   # :62-63
            raise ValueError(f"Expected {self.num_layers} predictions, got {len(predictions)}")
        empty = [p is None or np.size(p) == 0 or np.shape(p)[0] == 0 for p in predictions]
        return not any(empty)

    def decode_predictions(self, predictions) -> np.ndarray:
        """(B, cells, 5 + C) float64: ``[x, y, w, h, score, class probabilities]``
        (reference :48-183)."""
        if not self._usable(predictions):
            # the reference silently drops empty scales (:71-72); a head with a missing
            # scale is not something the CUDA path decodes
            return np.zeros((0, 0, 5 + self.num_classes), dtype="float32")
        return engine.decode_dense(predictions, self.anchors, self.num_classes, self.input_shape,
                                   None, self.use_softmax, self.rescore_confidence)

    def correct_boxes(self, predictions: np.ndarray, image_shape, model_image_size) -> np.ndarray:
        """Letterbox un-correction of an already decoded tensor (reference :185-235):
        pure per-element affine map with float32 constants, evaluated in float64."""
        model = np.array(model_image_size, dtype="float32")
        image = np.array(image_shape, dtype="float32")
        fitted = np.round(image * np.min(model / image))
        offset = ((model - fitted) / 2.0 / model)[..., ::-1]
        scale = (model / fitted)[..., ::-1]
        xy = (predictions[..., 0:2] - offset) * scale
        wh = predictions[..., 2:4] * scale
        xy = (xy - wh / 2.0) * image[..., ::-1]
        wh = wh * image[..., ::-1]
        return np.concatenate([xy, wh, predictions[..., 4:5], predictions[..., 5:]], axis=-1)

    def handle_predictions(self, predictions: np.ndarray, image_shape, max_boxes: int = 100,
                           confidence: float = 0.1, nms_threshold: float = 0.5,
                           use_iol: bool = True, nms_method: str = "diou", use_wbf: bool = False):
        """Threshold + NMS + top-k on a decoded, corrected tensor (reference :237-345).
        Candidate selection is indexing glue; the NMS itself runs on the GPU."""
        scores_all = predictions[..., 4]
        classes_all = np.argmax(predictions[..., 5:], axis=-1)
        pos = np.where(scores_all >= confidence)
        if len(pos[0]) == 0:
            return _EMPTY()
        boxes, classes, scores = predictions[..., 0:4][pos], classes_all[pos], scores_all[pos]
        if use_wbf:                                                          # :281-287
            n_boxes, n_classes, n_scores = WeightedBoxesFusion(iou_thr=nms_threshold).fuse_boxes(
                [boxes], [classes], [scores], image_shape)
            if not n_boxes:
                return _EMPTY()
            boxes, classes, scores = n_boxes[0], n_classes[0].astype("int32"), n_scores[0]
            if len(boxes) <= max_boxes:
                return boxes, classes, scores
            top = np.argsort(-scores, kind="stable")[:max_boxes]
            return boxes[top], classes[top], scores[top]
        nms = {"diou": DIoUNMS, "cluster": ClusterNMS, "standard": StandardNMS,
               "soft": SoftNMS}.get(nms_method, NMS)
        nms = nms() if nms is SoftNMS else nms(use_iol=use_iol)
        n_boxes, n_classes, n_scores = nms.apply_nms(boxes, classes, scores, nms_threshold, confidence)
        if not n_boxes:
            return _EMPTY()
        boxes = np.concatenate(n_boxes)
        classes = np.concatenate(n_classes).astype("int32")
        scores = np.concatenate(n_scores)
        if len(boxes) <= max_boxes:                                          # :336-337
            return boxes, classes, scores
        top = np.argsort(-scores, kind="stable")[:max_boxes]                 # :340-345
        return boxes[top], classes[top], scores[top]

    # ---- the hot path ------------------------------------------------------------
    def postprocess(self, multigriddet_outputs, image_shape, model_image_size,
                    max_boxes: int = 100, confidence: float = 0.1, nms_threshold: float = 0.5,
                    use_iol: bool = True, nms_method: str = "diou", use_wbf: bool = False,
                    return_xyxy: bool = True):
        """Complete postprocessing of ONE image (reference :347-395).

        Returns ``(boxes, classes, scores)``: int32 (K, 4) xyxy (or float64 xywh when
        ``return_xyxy=False``), int32 (K,), float64 (K,); three empty arrays when
        nothing passes.  ``nms_method``: 'diou', 'cluster', 'soft' or 'standard'
        ('standard' raises ``NotImplementedError`` in the reference; here it is IoU greedy NMS).
        """
        if use_wbf:
            nms_method = "wbf"                      # fusion with iou_thr = nms_threshold, :283
        elif nms_method not in ("diou", "cluster", "standard", "soft"):
            raise NotImplementedError(f"nms_method={nms_method!r} is not part of the CUDA path")
        if not self._usable(multigriddet_outputs):
            return _EMPTY()
        if int(np.shape(multigriddet_outputs[0])[0]) != 1:
            raise ValueError("postprocess() takes one image (batch 1) like every caller of the "
                             "reference; use postprocess_batch() for B > 1")
        out = self.postprocess_batch(multigriddet_outputs, [image_shape], model_image_size,
                                     max_boxes, confidence, nms_threshold, nms_method,
                                     return_xyxy=return_xyxy)
        return out[0]

    def postprocess_batch(self, multigriddet_outputs, image_shapes, model_image_size,
                          max_boxes: int = 100, confidence: float = 0.1,
                          nms_threshold: float = 0.5, nms_method: str = "diou",
                          per_class: bool = False, return_xyxy: bool = True):
        """B independent ``postprocess`` calls in one launch; returns a list of B
        ``(boxes, classes, scores)`` triples."""
        # the reference normalises wh by self.input_shape (:163) and un-letterboxes with
        # model_image_size (:205-216); the kernel has one model size, so they must agree
        if tuple(int(v) for v in model_image_size) != tuple(int(v) for v in self.input_shape):
            raise ValueError(f"model_image_size {tuple(model_image_size)} differs from the decoder's "
                             f"input_shape {tuple(self.input_shape)}: construct the decoder with "
                             "input_shape=model_image_size")
        want = ("boxes_xyxy" if return_xyxy else "boxes_xywh", "scores", "classes")
        det = engine.decode_nms(multigriddet_outputs, image_shapes, model_image_size,
                                self.anchors, self.num_classes, max_boxes, confidence,
                                nms_threshold, nms_method, per_class, self.use_softmax,
                                self.rescore_confidence, want=want)
        to_np = (lambda t: t.cpu().numpy()) if engine._is_torch(det["counts"]) else (lambda t: t)
        counts = to_np(det["counts"])
        boxes = to_np(det[want[0]])
        scores = to_np(det["scores"])
        classes = to_np(det["classes"])
        res = []
        for b, k in enumerate(counts):
            if k == 0:
                res.append(_EMPTY())
            else:
                res.append((boxes[b, :k].copy(), classes[b, :k].copy(), scores[b, :k].copy()))
        return res

    # kept for callers that post-process their own xywh boxes (reference :397-422)
    def _convert_to_xyxy(self, boxes: np.ndarray, image_shape) -> np.ndarray:
        b = np.array(boxes, dtype=np.float64, copy=True)
        b[:, 2] = boxes[:, 0] + boxes[:, 2]
        b[:, 3] = boxes[:, 1] + boxes[:, 3]
        h, w = image_shape[0], image_shape[1]
        b[:, 0::2] = np.clip(b[:, 0::2], 0, w)
        b[:, 1::2] = np.clip(b[:, 1::2], 0, h)
        return np.floor(b + 0.5).astype("int32")
