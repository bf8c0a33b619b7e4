"""Target-encoding half of the grid path (reference ``multigriddet.data``).

Only the hot-path entry points exist here; the image pipeline (generators,
augmentation, preprocessing) stays in the reference.
"""
from .generators import (get_anchor_mask, preprocess_true_boxes,
                         tf_preprocess_true_boxes)
from .target_encoding import MultiGridConfig, MultiGridTargetEncoder

__all__ = ["preprocess_true_boxes", "tf_preprocess_true_boxes", "get_anchor_mask",
           "MultiGridConfig", "MultiGridTargetEncoder"]
