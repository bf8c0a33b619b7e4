"""Names of the reference's ``multigriddet/data/target_encoding.py``, routed to the
CUDA encoder.

The reference module is orphaned and unit-inconsistent (normalised box sizes against
pixel anchors); it is not an oracle.  The classes below keep the
import surface (``MultiGridConfig``, ``MultiGridTargetEncoder``, the compat
``preprocess_true_boxes``) and compute the *generator* semantics
(``generators.py:3393``), i.e. boxes are pixel ``[x1, y1, x2, y2, class]``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import generators


@dataclass
class MultiGridConfig:
    """Same fields as the reference dataclass (target_encoding.py:14-24)."""
    input_shape: Tuple[int, int] = (608, 608)
    num_classes: int = 80
    anchors: Optional[List[np.ndarray]] = None
    num_layers: int = 3
    grid_assignment: str = "3x3"
    iou_threshold: float = 0.2
    multi_anchor_assign: bool = False
    max_boxes: int = 100


class MultiGridTargetEncoder:
    """IoL anchor matching + 3x3 dense grid assignment on the GPU."""

    def __init__(self, config: MultiGridConfig):
        self.config = config
        if config.anchors is None:       # target_encoding.py:38-43 defaults
            config.anchors = [np.array([[10, 13], [16, 30], [33, 23]], dtype=np.float32),
                              np.array([[30, 61], [62, 45], [59, 119]], dtype=np.float32),
                              np.array([[116, 90], [156, 198], [373, 326]], dtype=np.float32)]
        self.anchors = config.anchors
        self.num_layers = len(self.anchors)
        strides = (32, 16, 8, 4, 2)
        self.grid_shapes = [(config.input_shape[0] // strides[l], config.input_shape[1] // strides[l])
                            for l in range(self.num_layers)]

    def encode_batch_targets(self, batch_boxes: List[np.ndarray]) -> List[np.ndarray]:
        """``batch_boxes``: list of (n_i, 5) pixel boxes -> list of L y_true arrays."""
        n_max = max([len(b) for b in batch_boxes] + [1])
        dense = np.zeros((len(batch_boxes), n_max, 5), dtype=np.float32)
        for i, b in enumerate(batch_boxes):
            b = np.asarray(b, dtype=np.float32).reshape(-1, 5)
            dense[i, :len(b)] = b
        return generators.preprocess_true_boxes(dense, self.config.input_shape, self.anchors,
                                                self.config.num_classes,
                                                self.config.multi_anchor_assign, self.grid_shapes)

    def encode_targets(self, boxes: np.ndarray) -> List[np.ndarray]:
        """One image: (n, 5) pixel boxes -> list of L (Gh, Gw, 5+A+C) arrays."""
        return [y[0] for y in self.encode_batch_targets([boxes])]


def preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes,
                          multi_anchor_assign=False, iou_threshold=0.2):
    """Compat entry point (target_encoding.py:347-377)."""
    return generators.preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes,
                                            multi_anchor_assign, None, iou_threshold)
