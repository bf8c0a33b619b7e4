"""multigriddet_b200 -- B200-native (sm_100a) detection-head grid path for MultiGridDet.

Drop-in replacements for the two hot-path modules of solufast-cvprojects/multigriddet,
backed by hand-written CUDA kernels behind a C-ABI shared library (``libmgd.so``,
``include/mgd.h``):

* ``multigriddet_b200.data``         -- multi-grid ``y_true`` target encoding
  (reference ``multigriddet/data/generators.py:3393``, ``data/target_encoding.py``)
* ``multigriddet_b200.postprocess``  -- dense head decode, threshold, NMS
  (reference ``multigriddet/postprocess/multigrid_decode.py``, ``nms.py``,
  ``gpu_postprocess.py``)
* ``multigriddet_b200.engine``       -- array-level API (NumPy / torch CUDA / DLPack)
* ``multigriddet_b200.sharding``     -- image-sharded multi-GPU driver

Everything else in the reference (backbone, loss, augmentation, trainer, mAP) is out
of scope and stays where it is.  There is no CPU fallback: importing works anywhere,
computing needs a B200.
"""

__version__ = "0.1.0"

__all__ = ["engine", "data", "postprocess", "synth", "sharding", "__version__"]
