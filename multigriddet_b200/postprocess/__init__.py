"""Decode / NMS half of the grid path (reference ``multigriddet.postprocess``)."""
from .multigrid_decode import MultiGridDecoder
from .nms import (NMS, ClusterNMS, DIoUNMS, SoftNMS, StandardNMS, fast_cluster_nms_boxes,
                  nms_boxes)
from .wbf import WeightedBoxesFusion, weighted_boxes_fusion
from .gpu_postprocess import multigriddet_postprocess_gpu

__all__ = ["MultiGridDecoder", "NMS", "StandardNMS", "DIoUNMS", "SoftNMS", "ClusterNMS",
           "nms_boxes", "fast_cluster_nms_boxes", "multigriddet_postprocess_gpu",
           "WeightedBoxesFusion", "weighted_boxes_fusion"]
