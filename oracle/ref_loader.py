"""Load the REAL reference implementation of the hot path from ``/root/reference``.

Test infrastructure (see ``oracle/__init__.py``).  Only usable in the build
container: ``/root/reference`` does not exist on the GPU box, so nothing that
runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this module.
It is used by ``oracle/gen_golden.py`` (to create ``tests/golden/*.npz``) and by
the ``-m "not gpu"`` tests that pin ``oracle/mgd_oracle.py`` against the
reference when the reference tree is present.

How the reference is loaded (it cannot be imported as a package because
``multigriddet/__init__.py`` pulls in TensorFlow, which is not installed):

* encoder: the four pure-NumPy ``FunctionDef`` nodes ``get_anchor_mask``,
  ``iol_common_center``, ``best_fit_and_layer`` and ``preprocess_true_boxes``
  are cut out of ``multigriddet/data/generators.py`` (:2473-2544, :3393-3473)
  with ``ast`` and exec'd with only ``np`` in scope.  No reference text is
  copied into this repo; the source is read where it lies at run time.
* decoder / NMS: ``multigriddet/postprocess/{nms,wbf,multigrid_decode}.py`` are
  loaded by file path under a stub ``tensorflow`` module
  (``multigrid_decode.py:9`` imports TF but never uses it).

``np.argsort`` on AVX-512/AVX2 hosts is not stable; the reference relies on it
for anchor choice under rounded-IoL ties (``generators.py:2530``).  Call
``pin_numpy_env()`` *before* NumPy is first imported to get the portable
(lowest-index-first) behaviour the oracle and the CUDA kernels implement.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("MGD_REFERENCE_ROOT", "/root/reference")

NUMPY_PIN = ("AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL "
             "AVX512_SPR AVX2 FMA3")


def pin_numpy_env() -> None:
    """Disable NumPy's AVX dispatch (stable small argsort, libm transcendentals).

    Must run before the first ``import numpy`` of the process to take effect.
    """
    os.environ.setdefault("NPY_DISABLE_CPU_FEATURES", NUMPY_PIN)


def numpy_is_pinned() -> bool:
    return os.environ.get("NPY_DISABLE_CPU_FEATURES", "") == NUMPY_PIN


STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "multigriddet", "data", "generators.py"))


def staged() -> bool:
    """True when ``oracle/build_ref.py`` has staged the reference's hot-path sources under
    ``oracle/_ref/`` (they travel to the GPU box; ``/root/reference`` does not)."""
    return os.path.isfile(os.path.join(STAGED_ROOT, "MANIFEST.json"))


def load_staged_encoder():
    """``preprocess_true_boxes`` from the staged verbatim source segments."""
    if "enc_staged" in _cache:
        return _cache["enc_staged"]
    spec = importlib.util.spec_from_file_location(
        "_mgd_ref_staged_encoder", os.path.join(STAGED_ROOT, "encoder_functions.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    raw = module.preprocess_true_boxes

    def preprocess_true_boxes(*a, **k):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            return raw(*a, **k)

    _cache["enc_staged"] = preprocess_true_boxes
    return preprocess_true_boxes


def load_tf_encoder():
    """The reference's TensorFlow encoder ``tf_preprocess_true_boxes`` (generators.py:2696-3390,
    the default training path) executed over ``oracle/tf_shim.py``: the function's own source,
    cut out with ``ast`` where it lies, with a NumPy stand-in answering its ``tf.*`` calls
    (TensorFlow is not installed here).  Returns ``f(true_boxes, input_shape, anchors,
    num_classes, grid_shapes) -> list of float32 arrays``."""
    if "tf_enc" in _cache:
        return _cache["tf_enc"]
    import typing
    import numpy as np
    from . import tf_shim
    path = os.path.join(REFERENCE_ROOT, "multigriddet", "data", "generators.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    picked = [n for n in tree.body
              if isinstance(n, ast.FunctionDef) and n.name == "tf_preprocess_true_boxes"]
    if len(picked) != 1:
        raise RuntimeError("tf_preprocess_true_boxes not found in " + path)
    scope = {"tf": tf_shim, "np": np, "Tuple": typing.Tuple, "List": typing.List,
             "Optional": typing.Optional, "Dict": typing.Dict, "Union": typing.Union}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), scope)
    raw = scope["tf_preprocess_true_boxes"]

    def tf_preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes, grid_shapes):
        tb = np.asarray(true_boxes, dtype=np.float32)
        anc = [np.asarray(a, dtype=np.float32) for a in anchors]
        out = raw(tb, tuple(int(v) for v in input_shape), anc, int(num_classes), False,
                  [tuple(int(v) for v in g) for g in grid_shapes])
        return [np.asarray(y, dtype=np.float32) for y in out]

    _cache["tf_enc"] = tf_preprocess_true_boxes
    return tf_preprocess_true_boxes


_ENCODER_NAMES = ("get_anchor_mask", "iol_common_center", "best_fit_and_layer",
                  "preprocess_true_boxes")
_cache: dict = {}


def load_encoder():
    """Return the reference's ``preprocess_true_boxes`` (NumPy encoder)."""
    if "enc" in _cache:
        return _cache["enc"]
    import numpy as np
    path = os.path.join(REFERENCE_ROOT, "multigriddet", "data", "generators.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    picked = [n for n in tree.body
              if isinstance(n, ast.FunctionDef) and n.name in _ENCODER_NAMES]
    if len(picked) != len(_ENCODER_NAMES):
        raise RuntimeError("reference encoder functions not found in " + path)
    mod = ast.Module(body=picked, type_ignores=[])
    scope = {"np": np}
    exec(compile(mod, path, "exec"), scope)
    raw = scope["preprocess_true_boxes"]

    def preprocess_true_boxes(*a, **k):
        # generators.py:3441 does int(<1-element array>), deprecated in NumPy 2
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            return raw(*a, **k)

    preprocess_true_boxes.scope = scope
    _cache["enc"] = preprocess_true_boxes
    return preprocess_true_boxes


def load_postprocess(staged_copy: bool = False):
    """Return a namespace with the reference ``MultiGridDecoder`` and NMS classes
    (``staged_copy``: from ``oracle/_ref`` instead of ``/root/reference``)."""
    key = "post_staged" if staged_copy else "post"
    if key in _cache:
        return _cache[key]
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    pkg_root = os.path.join(STAGED_ROOT if staged_copy else REFERENCE_ROOT, "multigriddet")
    names = {}
    for pkg, sub in (("_mgd_ref", pkg_root),
                     ("_mgd_ref.postprocess", os.path.join(pkg_root, "postprocess"))):
        m = types.ModuleType(pkg)
        m.__path__ = [sub]
        sys.modules[pkg] = m
    for name in ("nms", "wbf", "multigrid_decode"):
        full = "_mgd_ref.postprocess." + name
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(pkg_root, "postprocess", name + ".py"))
        module = importlib.util.module_from_spec(spec)
        sys.modules[full] = module
        spec.loader.exec_module(module)
        names[name] = module
    ns = types.SimpleNamespace(
        MultiGridDecoder=names["multigrid_decode"].MultiGridDecoder,
        NMS=names["nms"].NMS, StandardNMS=names["nms"].StandardNMS,
        DIoUNMS=names["nms"].DIoUNMS, SoftNMS=names["nms"].SoftNMS,
        ClusterNMS=names["nms"].ClusterNMS, nms_boxes=names["nms"].nms_boxes,
        fast_cluster_nms_boxes=names["nms"].fast_cluster_nms_boxes,
        WeightedBoxesFusion=names["wbf"].WeightedBoxesFusion)
    _cache[key] = ns
    return ns


def load_metrics():
    """Return the reference's ``multigriddet/evaluation/metrics.py`` module (pure NumPy /
    Python; its ``..utils.boxes`` import needs the stub ``tensorflow`` only at import)."""
    if "metrics" in _cache:
        return _cache["metrics"]
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    pkg_root = os.path.join(REFERENCE_ROOT, "multigriddet")
    for pkg, sub in (("_mgd_ref", pkg_root), ("_mgd_ref.utils", os.path.join(pkg_root, "utils")),
                     ("_mgd_ref.evaluation", os.path.join(pkg_root, "evaluation"))):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [sub]
            sys.modules[pkg] = m
    out = None
    for full, rel in (("_mgd_ref.utils.boxes", ("utils", "boxes.py")),
                      ("_mgd_ref.evaluation.metrics", ("evaluation", "metrics.py"))):
        spec = importlib.util.spec_from_file_location(full, os.path.join(pkg_root, *rel))
        module = importlib.util.module_from_spec(spec)
        sys.modules[full] = module
        spec.loader.exec_module(module)
        out = module
    _cache["metrics"] = out
    return out


def load_box_transforms():
    """The reference's ``reshape_boxes`` and ``merge_mosaic_bboxes`` (pure NumPy;
    ``augmentation.py`` itself imports cv2 / PIL / imgaug, so the two ``FunctionDef`` nodes
    are cut out with ``ast`` like the encoder).  ``reshape_boxes`` shuffles its rows with
    ``np.random.shuffle`` (:146); the returned wrapper runs it with the shuffle disabled so
    results are comparable row by row (and on a copy: the reference writes in place)."""
    if "boxes" in _cache:
        return _cache["boxes"]
    import numpy as np
    path = os.path.join(REFERENCE_ROOT, "multigriddet", "data", "augmentation.py")
    with open(path, "r") as fh, warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)       # a docstring of that file has "\ "
        tree = ast.parse(fh.read(), filename=path)
    names = ("reshape_boxes", "merge_mosaic_bboxes")
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    if len(picked) != len(names):
        raise RuntimeError("reference box transforms not found in " + path)

    class _NoShuffleRandom:
        @staticmethod
        def shuffle(_):
            return None

    class _Np:
        """NumPy with ``random.shuffle`` turned into a no-op."""
        random = _NoShuffleRandom()

        def __getattr__(self, name):
            return getattr(np, name)

    scope = {"np": _Np()}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), scope)
    raw = scope["reshape_boxes"]

    def reshape_boxes(boxes, *a, **k):
        return raw(np.array(boxes), *a, **k)

    ns = types.SimpleNamespace(reshape_boxes=reshape_boxes,
                               merge_mosaic_bboxes=scope["merge_mosaic_bboxes"])
    _cache["boxes"] = ns
    return ns
