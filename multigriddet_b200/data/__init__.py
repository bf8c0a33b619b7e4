"""Target-encoding half of the grid path (reference ``multigriddet.data``).

Only the hot-path entry points exist here (plus the two box-side functions of the
augmentation module that feed it); the image pipeline (generators, image augmentation,
preprocessing) stays in the reference.
"""
from .augmentation import merge_mosaic_bboxes, reshape_boxes
from .generators import (expand_box_capacity, get_anchor_mask, letterbox_boxes,
                         preprocess_true_boxes, tf_preprocess_true_boxes)
from .target_encoding import MultiGridConfig, MultiGridTargetEncoder

__all__ = ["preprocess_true_boxes", "tf_preprocess_true_boxes", "get_anchor_mask",
           "MultiGridConfig", "MultiGridTargetEncoder", "reshape_boxes", "merge_mosaic_bboxes",
           "letterbox_boxes", "expand_box_capacity"]
