"""Drop-in for the reference's ``multigriddet/postprocess/gpu_postprocess.py``.

The reference function is an uncalled TensorFlow restatement whose only
device-specific step is ``tf.image.combined_non_max_suppression`` (per-class NMS,
third-party, no test pins it).  This module keeps the signature and the padded
return layout and routes to the CUDA path with the *NumPy path's* decode semantics
(cell offsets ``[col, row]``, rounded letterbox size) and per-class greedy IoU NMS as
defined in DESIGN.md; parity against the TF op is unpinned.
"""
from __future__ import annotations

import numpy as np

from .. import engine


def multigriddet_postprocess_gpu(multigriddet_outputs, image_shapes, anchors, num_classes,
                                 model_image_size, max_boxes=500, confidence=0.001,
                                 nms_threshold=0.45, rescore_confidence=True, use_iol=True):
    """Batched decode + per-class NMS.  Returns ``(boxes, scores, classes, valid)`` in
    the order the reference actually returns them (gpu_postprocess.py:282):
    (B, max, 4) float32 ``[x1, y1, x2, y2]`` pixels, (B, max) float32 scores,
    (B, max) float32 class ids, (B,) int32 valid counts; rows past ``valid`` are 0."""
    del use_iol
    det = engine.decode_nms(multigriddet_outputs, image_shapes, model_image_size, anchors,
                            num_classes, max_boxes, confidence, nms_threshold, "standard",
                            per_class=True, use_softmax=True,
                            rescore_confidence=rescore_confidence,
                            want=("boxes_xywh", "scores", "classes"))
    if engine._is_torch(det["counts"]):
        import torch
        xywh = det["boxes_xywh"]
        boxes = torch.cat([xywh[..., :2], xywh[..., :2] + xywh[..., 2:]], -1).to(torch.float32)
        pad = (det["classes"] < 0).unsqueeze(-1)
        boxes = torch.where(pad, torch.zeros_like(boxes), boxes)
        classes = det["classes"].clamp(min=0).to(torch.float32)
        return boxes, det["scores"].to(torch.float32), classes, det["counts"]
    xywh = det["boxes_xywh"]
    boxes = np.concatenate([xywh[..., :2], xywh[..., :2] + xywh[..., 2:]], -1).astype(np.float32)
    boxes[det["classes"] < 0] = 0
    classes = np.maximum(det["classes"], 0).astype(np.float32)
    return boxes, det["scores"].astype(np.float32), classes, det["counts"]
