"""Loss-side consumer of ``y_true`` on the grid path (reference ``multigriddet.losses``):
only the forward-only ignore-mask computation lives here; the loss itself stays in the
reference's TensorFlow graph."""
from .ignore_mask import compute_ignore_mask, compute_ignore_masks

__all__ = ["compute_ignore_mask", "compute_ignore_masks"]
