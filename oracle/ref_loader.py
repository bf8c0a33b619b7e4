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


def load_tf_ignore_mask():
    """The reference's loss-side ignore mask -- ``MultiGridLoss._compute_ignore_mask`` with
    ``_compute_iou_batch`` (losses/multigrid_loss.py:445-703) -- executed over
    ``oracle/tf_shim.py``: the two methods' own source, cut out of the class with ``ast`` and
    bound to a bare object that carries the three attributes they read (``input_shape``,
    ``ignore_thresh``, ``eps`` = ``K.epsilon()``, :122-123, :173).  The returned function takes
    one layer the way the caller does (:285-320): ``f(y_pred_layer, y_true_layer, anchors_layer,
    input_shape, ignore_thresh) -> (ignore_mask, assigned_anchor_iou, max_iou_map)``."""
    if "tf_ignore" in _cache:
        return _cache["tf_ignore"]
    import typing
    import numpy as np
    from . import tf_shim
    path = os.path.join(REFERENCE_ROOT, "multigriddet", "losses", "multigrid_loss.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    wanted = ("_compute_iou_batch", "_compute_ignore_mask")
    picked = [m for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "MultiGridLoss"
              for m in c.body if isinstance(m, ast.FunctionDef) and m.name in wanted]
    if len(picked) != len(wanted):
        raise RuntimeError("ignore-mask methods not found in " + path)
    holder = ast.ClassDef(name="_RefLoss", bases=[], keywords=[], body=picked, decorator_list=[])
    if "type_params" in ast.ClassDef._fields:
        holder.type_params = []
    mod = ast.fix_missing_locations(ast.Module(body=[holder], type_ignores=[]))
    scope = {"tf": tf_shim, "K": tf_shim.keras.backend, "np": np, "Tuple": typing.Tuple,
             "List": typing.List, "Optional": typing.Optional}
    exec(compile(mod, path, "exec"), scope)
    cls = scope["_RefLoss"]

    def ignore_mask_layer(y_pred, y_true, anchors, input_shape, ignore_thresh=0.5):
        y_pred = np.asarray(y_pred, dtype=np.float32)
        y_true = np.asarray(y_true, dtype=np.float32)
        obj = cls()
        obj.input_shape = tuple(int(v) for v in input_shape)
        obj.ignore_thresh = ignore_thresh
        obj.eps = tf_shim.keras.backend.epsilon()
        object_mask = (y_true[..., 4:5] > 0.5).astype(np.float32)            # :305
        grid_shape = (np.int32(y_pred.shape[1]), np.int32(y_pred.shape[2]))  # :308
        out = obj._compute_ignore_mask(y_pred[..., 0:2], y_pred[..., 2:4], y_true[..., 0:2],
                                       y_true[..., 2:4], np.asarray(anchors), object_mask, y_true,
                                       grid_shape)
        return tuple(np.asarray(o, dtype=np.float32) for o in out)

    _cache["tf_ignore"] = ignore_mask_layer
    return ignore_mask_layer


def load_tf_box_prestep():
    """The box side of the reference's tf.data pipeline (``build_tf_dataset``) executed over
    ``oracle/tf_shim.py``: ``tf_letterbox_resize`` (:167-209), ``tf_random_horizontal_flip``
    (:227-256) and the two closures ``_preprocess_image_and_boxes`` (:1859-1958) and
    ``_expand_box_capacity`` (:1983-2034), each function's own source cut out with ``ast``.
    The closures' free variables (``self``, ``has_multiscale``, ``input_shape_list_tf``,
    ``input_shape_base_tf``) are supplied per call; images are zero arrays of the right shape
    (the box arithmetic only reads their shapes); ``self.augment`` is False for the transform
    (the crop / colour / rotate augmentations between the two box steps are not on the path) and
    the flip is called on its own with the coin forced; ``padded_batch`` (:1963-1976), a
    tf.data method, is the zero-padding to ``max_boxes_per_image`` rows it documents.

    Returns ``f(boxes (n, 5), src_hw, input_shape, max_boxes_per_image, expansion,
    multiscale_shape=None, hflip=False) -> (max_boxes_per_image * expansion, 5) float32``."""
    if "tf_boxes" in _cache:
        return _cache["tf_boxes"]
    import types as _types
    import typing
    import numpy as np
    from . import tf_shim
    path = os.path.join(REFERENCE_ROOT, "multigriddet", "data", "generators.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    top = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    nested = {}                  # the closures of build_tf_dataset: the FIRST definitions in the file
    for n in ast.walk(tree):     # (a later method defines closures of the same names)
        if isinstance(n, ast.FunctionDef) and n.name in ("_preprocess_image_and_boxes", "_expand_box_capacity"):
            if n.name not in nested or n.lineno < nested[n.name].lineno:
                nested[n.name] = n
    names = ("tf_letterbox_resize", "tf_random_horizontal_flip", "tf_normalize_image")
    if any(k not in top for k in names) or len(nested) != 2:
        raise RuntimeError("tf.data box functions not found in " + path)
    body = [top[k] for k in names] + [nested["_preprocess_image_and_boxes"], nested["_expand_box_capacity"]]
    scope = {"tf": tf_shim, "np": np, "Tuple": typing.Tuple, "List": typing.List,
             "Optional": typing.Optional, "Dict": typing.Dict, "Union": typing.Union}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), scope)

    def run(boxes, src_hw, input_shape, max_boxes_per_image, expansion=1, multiscale_shape=None,
            hflip=False):
        input_shape = tuple(int(v) for v in input_shape)
        me = _types.SimpleNamespace(
            input_shape=input_shape, augment=False, rescale_interval=10 if multiscale_shape else -1,
            max_boxes_per_image=int(max_boxes_per_image),
            enhance_augment="mosaic" if expansion in (4, 8) else None,
            mosaic_prob=1.0 if expansion in (4, 8) else 0.0,
            mixup_prob=1.0 if expansion in (2, 8) else 0.0)
        scope["self"] = me
        scope["has_multiscale"] = multiscale_shape is not None                      # :1849-1857
        scope["input_shape_list_tf"] = (np.array([multiscale_shape], dtype=np.int32)
                                        if multiscale_shape is not None else None)
        scope["input_shape_base_tf"] = np.array(input_shape, dtype=np.int32)
        tf_shim.random.forced[:] = [0] if multiscale_shape is not None else []       # the scale index drawn at :1868
        image = np.zeros((int(src_hw[0]), int(src_hw[1]), 3), dtype=np.uint8)
        bx = np.asarray(boxes, dtype=np.float32).reshape(-1, 5)
        image_out, bx = scope["_preprocess_image_and_boxes"](image, bx, None)
        tf_shim.random.forced[:] = [0.75 if hflip else 0.25]                        # the coin drawn at :241
        _, bx = scope["tf_random_horizontal_flip"](image_out, bx)
        bx = np.asarray(bx, dtype=np.float32)
        if bx.shape[0] > me.max_boxes_per_image:
            raise ValueError("padded_batch raises on a component longer than its padded shape")
        dense = np.zeros((1, me.max_boxes_per_image, 5), dtype=np.float32)          # padded_batch :1963-1976
        dense[0, :bx.shape[0]] = bx
        _, expanded = scope["_expand_box_capacity"](np.zeros((1,) + input_shape + (3,), np.float32), dense)
        return np.asarray(expanded, dtype=np.float32)[0]

    _cache["tf_boxes"] = run
    return run


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
