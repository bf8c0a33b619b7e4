"""A NumPy stand-in for the TensorFlow ops the reference's TF code of the hot path calls.

Test infrastructure (see ``oracle/__init__.py``).  TensorFlow is not installed in this image,
so the reference's TensorFlow functions cannot be run as they are.  What CAN be run is their
own source, statement by statement, with every ``tf.*`` call answered by the NumPy function
below that implements the op's documented semantics on float32 / int32 / int64 / bool arrays:
``oracle/ref_loader.load_tf_encoder()`` cuts ``tf_preprocess_true_boxes`` out of
``multigriddet/data/generators.py`` (:2696-3390) with ``ast`` and executes it with this module
bound to the name ``tf``; ``load_tf_ignore_mask()`` does the same with the two methods
``MultiGridLoss._compute_iou_batch`` / ``_compute_ignore_mask``
(``multigriddet/losses/multigrid_loss.py:445-703``; ``K`` = ``keras.backend`` below).  The control flow, the op order, the indexing and the constants are
then the reference's; what is assumed is only what each primitive op does:

* element-wise arithmetic / comparison in the operands' dtype (float32 stays float32; a
  Python scalar takes the tensor's dtype, like TF's constant conversion and NumPy 2's weak
  scalars), ``tf.cast`` float -> int truncates toward zero;
* ``tf.argmax`` returns the first maximum; ``tf.where(cond)`` lists true positions in
  row-major order (int64); ``tf.gather`` / ``tf.gather_nd`` / slicing index like NumPy;
* ``tf.tensor_scatter_nd_update`` applies the updates in order on the CPU (a later update to
  the same position wins; TF documents the order as undefined in general and sequential on
  CPU), ``tf.math.unsorted_segment_sum`` sums by segment id, ``tf.one_hot`` yields a zero row
  for an index outside [0, depth);
* ``tf.math.log`` / ``tf.exp`` / ``tf.nn.tanh`` / ``tf.nn.sigmoid`` on float32 are NumPy's float32
  functions (glibc's with NumPy's AVX dispatch off, tests/conftest.py; sigmoid = 1 / (1 + exp(-x)))
  -- TensorFlow's are Eigen's, so values downstream of them are only claimed to 1e-5;
* ``tf.map_fn`` applies the function to the elements in order and stacks the results;
* ``tf.debugging.*`` and ``tf.print`` are no-ops, ``tf.function`` is the identity,
  ``tf.cond(pred, a, b)`` calls ``a()`` or ``b()``.

Only the ops that function uses are provided; anything else raises AttributeError, so a new
op in the reference cannot be silently mis-modelled.
"""
from __future__ import annotations

import numpy as np

float32, float64 = np.float32, np.float64
int32, int64 = np.int32, np.int64
bool = np.bool_          # noqa: A001  (tf.bool)
Tensor = np.ndarray


def _a(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def function(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def constant(value, dtype=None):
    return np.array(value, dtype=dtype)


def cast(x, dtype):
    x = _a(x)
    if np.issubdtype(np.dtype(dtype), np.integer) and np.issubdtype(x.dtype, np.floating):
        return np.trunc(x).astype(dtype)
    return x.astype(dtype)


def shape(x):
    return np.array(_a(x).shape, dtype=np.int32)


def rank(x):
    return np.int32(_a(x).ndim)


def size(x):
    return np.int32(_a(x).size)


def zeros(shp, dtype=np.float32):
    return np.zeros([int(v) for v in shp], dtype=dtype)


def ones(shp, dtype=np.float32):
    return np.ones([int(v) for v in shp], dtype=dtype)


def reshape(x, shp):
    return _a(x).reshape([int(v) for v in np.asarray(shp).reshape(-1)])


def expand_dims(x, axis):
    return np.expand_dims(_a(x), int(axis))


def squeeze(x, axis=None):
    return np.squeeze(_a(x), axis=axis)


def stack(values, axis=0):
    # TF converts Python scalars in the list to the dtype of the tensor elements
    dts = [np.asarray(v).dtype for v in values if isinstance(v, (np.ndarray, np.generic))]
    dt = dts[0] if dts else None
    return np.stack([_a(v, dtype=dt) if not isinstance(v, (np.ndarray, np.generic)) else _a(v)
                     for v in values], axis=axis)


def concat(values, axis=0):
    arrs = [_a(v) for v in values]
    # TF takes the dtype of the tensor operands; Python lists are converted to it
    tensor_dtypes = [a.dtype for v, a in zip(values, arrs) if isinstance(v, np.ndarray)]
    dt = tensor_dtypes[0] if tensor_dtypes else arrs[0].dtype
    return np.concatenate([a.astype(dt) for a in arrs], axis=axis)


def tile(x, multiples):
    return np.tile(_a(x), [int(v) for v in multiples])


def broadcast_to(x, shp):
    return np.broadcast_to(_a(x), [int(v) for v in np.asarray(shp).reshape(-1)]).copy()


def range(*args, dtype=None):     # noqa: A001  (tf.range)
    out = np.arange(*[int(a) for a in args])
    return out.astype(dtype if dtype is not None else np.int32)


def meshgrid(*xs, indexing="xy"):
    return [g.copy() for g in np.meshgrid(*[_a(x) for x in xs], indexing=indexing)]


def cumsum(x, axis=0):
    x = _a(x)
    return np.cumsum(x, axis=axis).astype(x.dtype)


def minimum(a, b):
    return np.minimum(a, b)


def maximum(a, b):
    a = _a(a)
    return np.maximum(a, _a(b, dtype=a.dtype) if not isinstance(b, np.ndarray) else b)


def reduce_max(x, axis=None):
    return np.max(_a(x), axis=axis)


def argmax(x, axis=None, output_type=np.int64):
    return np.argmax(_a(x), axis=axis).astype(output_type)


def equal(a, b):
    return np.equal(a, b)


def logical_and(a, b):
    return np.logical_and(a, b)


def logical_or(a, b):
    return np.logical_or(a, b)


def logical_not(a):
    return np.logical_not(a)


def where(condition, x=None, y=None):
    if x is None and y is None:
        return np.argwhere(_a(condition)).astype(np.int64)
    return np.where(condition, x, y)


def gather(params, indices, axis=0, batch_dims=0):
    if batch_dims:
        raise NotImplementedError("tf.gather with batch_dims")
    return np.take(_a(params), _a(indices), axis=axis)


def gather_nd(params, indices, batch_dims=0):
    if batch_dims:
        raise NotImplementedError("tf.gather_nd with batch_dims")
    params, indices = _a(params), _a(indices)
    return params[tuple(np.moveaxis(indices, -1, 0))]


def one_hot(indices, depth, dtype=np.float32):
    indices = _a(indices)
    depth = int(depth)
    out = np.zeros(indices.shape + (depth,), dtype=dtype)
    ok = (indices >= 0) & (indices < depth)
    pos = np.nonzero(ok)
    out[pos + (indices[ok],)] = 1
    return out


def tensor_scatter_nd_update(tensor, indices, updates):
    out = np.array(tensor, copy=True)
    indices, updates = _a(indices), _a(updates)
    for i in np.arange(indices.shape[0]):          # sequential: a later update wins (TF on CPU)
        out[tuple(indices[i])] = updates[i]
    return out


def exp(x):
    return np.exp(_a(x))


def reduce_sum(x, axis=None, keepdims=False):
    x = _a(x)
    return np.sum(x, axis=axis, keepdims=keepdims, dtype=x.dtype)


def tensordot(a, b, axes):
    a, b = _a(a), _a(b)
    return np.tensordot(a, b, axes=axes).astype(np.result_type(a.dtype, b.dtype))


def stop_gradient(x):
    return x


class TensorSpec:
    def __init__(self, shape=None, dtype=None):
        self.shape, self.dtype = shape, dtype


def map_fn(fn, elems, fn_output_signature=None, **_kw):
    return np.stack([_a(fn(e)) for e in _a(elems)], axis=0)


def cond(pred, true_fn, false_fn):
    return true_fn() if np.asarray(pred).item() else false_fn()


def print(*_a, **_k):          # noqa: A001  (tf.print)
    return None


class _Math:
    @staticmethod
    def log(x):
        return np.log(_a(x))

    @staticmethod
    def unsorted_segment_sum(data, segment_ids, num_segments):
        data, segment_ids = _a(data), _a(segment_ids)
        out = np.zeros((int(num_segments),) + data.shape[segment_ids.ndim:], dtype=data.dtype)
        np.add.at(out, segment_ids, data)
        return out


class _Debugging:
    @staticmethod
    def _noop(*_a, **_k):
        return None

    assert_equal = assert_shapes = assert_rank = assert_positive = assert_greater = _noop


class _NN:
    @staticmethod
    def tanh(x):
        return np.tanh(_a(x))

    @staticmethod
    def sigmoid(x):
        x = _a(x)
        one = x.dtype.type(1)
        return one / (one + np.exp(-x))


class _Backend:
    """tensorflow.keras.backend, as far as the loss code uses it"""

    @staticmethod
    def epsilon():
        return 1e-7

    shape = staticmethod(shape)
    maximum = staticmethod(maximum)
    minimum = staticmethod(minimum)

    @staticmethod
    def cast(x, dtype):
        return cast(x, np.dtype(dtype))


class _Keras:
    backend = _Backend()


class _Logging:
    info = "info"


class _V1:
    logging = _Logging()


class _Compat:
    v1 = _V1()


class _Image:
    """Image-space ops: the box-side functions only ever look at the SHAPES of the images they
    produce, so these return zero images of the documented output shape."""

    @staticmethod
    def resize(image, size, method="bilinear"):
        image = _a(image)
        return np.zeros((int(size[0]), int(size[1])) + image.shape[2:], dtype=np.float32)

    @staticmethod
    def pad_to_bounding_box(image, offset_height, offset_width, target_height, target_width):
        image = _a(image)
        return np.zeros((int(target_height), int(target_width)) + image.shape[2:], dtype=image.dtype)

    @staticmethod
    def flip_left_right(image):
        return image


class _Random:
    """tf.random.uniform answers from a queue the test fills (``forced``): the reference draws
    its coin / its scale index there, the tests decide them."""

    def __init__(self):
        self.forced = []

    def uniform(self, shape, minval=0, maxval=None, dtype=np.float32, **_k):
        if not self.forced:
            raise RuntimeError("tf_shim.random.uniform called with no forced value queued")
        return np.asarray(self.forced.pop(0), dtype=dtype)


math = _Math()
nn = _NN()
image = _Image()
random = _Random()
debugging = _Debugging()
keras = _Keras()
compat = _Compat()
