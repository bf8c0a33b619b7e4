"""Drop-in for the two box-side functions of the reference's
``multigriddet/data/augmentation.py`` (the image-side augmentations stay in the reference):
same names, arguments and return values; the arithmetic runs in ``libmgd.so``
(``mgd_reshape_boxes``, ``mgd_mosaic_merge_boxes``).  The batch forms in
``multigriddet_b200.engine`` (``reshape_boxes_batch``, ``mosaic_merge_boxes_batch``) take
device tensors and feed ``encode_targets`` without touching host memory.
"""
from __future__ import annotations

import numpy as np

from .. import engine


def reshape_boxes(boxes, src_shape, target_shape, padding_shape, offset, horizontal_flip=False,
                  vertical_flip=False):
    """Boxes of one image from ``src_shape`` (w, h) to the padded / flipped ``target_shape``
    image (reference augmentation.py:112-164).  Returns the surviving rows, (k, 5), in the
    dtype of ``boxes``.

    Two deliberate differences: the rows are NOT shuffled (the reference calls
    ``np.random.shuffle`` on them, :146 -- shuffle before calling if the order matters) and
    the caller's array is not modified in place.
    """
    b = np.asarray(boxes)
    if len(b) == 0:
        return boxes
    dtype = b.dtype
    work = b.astype(np.int32) if np.issubdtype(dtype, np.integer) else b.astype(np.float64)
    params = np.array([[src_shape[0], src_shape[1], target_shape[0], target_shape[1],
                        padding_shape[0], padding_shape[1], offset[0], offset[1],
                        int(bool(horizontal_flip)), int(bool(vertical_flip))]], dtype=np.int32)
    out, _, cnt = engine.reshape_boxes_batch(work[None], params, want_f32=False)
    return out[0, :int(cnt[0])].astype(dtype)


def merge_mosaic_bboxes(bboxes, crop_x, crop_y, image_size):
    """(4, N, 5) boxes of the four mosaic samples -> (N, 5) merged boxes, zero padded
    (reference augmentation.py:606-667)."""
    bboxes = np.asarray(bboxes)
    assert bboxes.shape[0] == 4, 'mosaic sample number should be 4'
    out, _, _ = engine.mosaic_merge_boxes_batch(bboxes, [[0, 1, 2, 3]], [[int(crop_x), int(crop_y)]],
                                                image_size, want_f32=False)
    return out[0]
