"""NumPy-backed stand-in for the slice of TensorFlow 2 (eager) that the reference's
dense-head routines touch.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installed in the build image, and the
reference (WD-Leong/CV-Lite-Object-Detection) does `import tensorflow as tf` at the top
of every module.  This stub lets `oracle/ref_loader.py` import the reference's library
modules *unmodified* from /root/reference so that `oracle/make_golden.py` can freeze
golden vectors from the reference's own source.  Nothing in the product path imports it.

Semantics that matter for parity and are reproduced here:
  * EagerTensor has `__array_priority__ = 100`, so `ndarray <op> Tensor` and
    `np.float32 <op> Tensor` dispatch to the Tensor's reflected operator.
  * In a binary op the non-Tensor operand (Python scalar, NumPy scalar, ndarray) is
    converted to the Tensor operand's dtype (`ops.convert_to_tensor(x, dtype=y.dtype)`),
    so a float64 ndarray minus a float32 Tensor is computed in float32.
  * Python floats / float lists become float32, Python ints become int32.
  * `.numpy()` returns a COPY (the reference mutates the result in place).
  * Mixing two Tensors of different dtypes raises, as TF does.
"""
import numpy as np

from . import keras  # noqa: F401  (tf.keras.*)

float16 = np.float16
float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_  # noqa: A001  (tf.bool)
uint8 = np.uint8

_py_bool = type(True)


def _default_dtype(x):
    """dtype TF would pick for a non-Tensor value with no dtype hint."""
    if isinstance(x, (np.ndarray, np.generic)):
        return x.dtype
    a = np.asarray(x)
    if a.dtype == np.float64:
        return np.dtype(np.float32)
    if a.dtype == np.int64:
        return np.dtype(np.int32)
    return a.dtype


class Tensor:
    __array_priority__ = 100

    def __init__(self, value, dtype=None):
        if isinstance(value, Tensor):
            value = value._v
        if dtype is None:
            dtype = _default_dtype(value)
        self._v = np.array(value, dtype=dtype)  # always a private copy

    # -- basic protocol ---------------------------------------------------
    @property
    def dtype(self):
        return self._v.dtype

    @property
    def shape(self):
        return tuple(self._v.shape)

    @property
    def ndim(self):
        return self._v.ndim

    def numpy(self):
        v = self._v.copy()
        return v[()] if v.ndim == 0 else v

    def __array__(self, dtype=None, copy=None):
        v = self._v
        return v.astype(dtype) if dtype is not None else v.copy()

    def __len__(self):
        if self._v.ndim == 0:
            raise TypeError("Scalar tensor has no len()")
        return self._v.shape[0]

    def __iter__(self):
        if self._v.ndim == 0:
            raise TypeError("Cannot iterate over a scalar tensor")
        return (Tensor(x) for x in self._v)

    def __getitem__(self, idx):
        if isinstance(idx, Tensor):
            idx = idx._v
        return Tensor(self._v[idx])

    def __int__(self):
        return int(self._v)

    def __float__(self):
        return float(self._v)

    def __index__(self):
        return int(self._v)

    def __bool__(self):
        return _py_bool(self._v)

    def __repr__(self):
        return "stub.Tensor(%r, dtype=%s)" % (self._v, self._v.dtype)

    __hash__ = object.__hash__

    # -- arithmetic ---------------------------------------------------------
    def _coerce(self, other):
        if isinstance(other, Tensor):
            if other.dtype != self.dtype:
                raise TypeError(
                    "stub tf: dtype mismatch %s vs %s" % (self.dtype, other.dtype))
            return other._v
        return np.asarray(other).astype(self.dtype)

    def _bin(self, other, fn, reflected=False):
        o = self._coerce(other)
        with np.errstate(all="ignore"):
            r = fn(o, self._v) if reflected else fn(self._v, o)
        return Tensor(r)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    def __pow__(self, o): return self._bin(o, np.power)
    def __rpow__(self, o): return self._bin(o, np.power, True)
    def __neg__(self): return Tensor(-self._v)
    def __abs__(self): return Tensor(np.abs(self._v))
    def __lt__(self, o): return self._bin(o, np.less)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __eq__(self, o): return self._bin(o, np.equal)
    def __ne__(self, o): return self._bin(o, np.not_equal)


def _t(x, dtype=None):
    return x if isinstance(x, Tensor) and dtype is None else Tensor(x, dtype)


def _pair(a, b):
    """convert two operands the way a TF binary op does; returns ndarrays."""
    if isinstance(a, Tensor):
        return a._v, a._coerce(b)
    if isinstance(b, Tensor):
        return b._coerce(a), b._v
    ta = Tensor(a)
    return ta._v, ta._coerce(b)


def convert_to_tensor(x, dtype=None):
    return Tensor(x, dtype)


def constant(x, dtype=None):
    return Tensor(x, dtype)


def cast(x, dtype):
    return Tensor(np.asarray(x._v if isinstance(x, Tensor) else x).astype(dtype))


def shape(x):
    return Tensor(np.array(np.shape(x._v if isinstance(x, Tensor) else x), dtype=np.int32))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    s, l, d = (float(v) if isinstance(v, (Tensor, float, np.floating)) else v
               for v in (start, limit, delta))
    out = np.arange(s, l, d)
    if dtype is None:
        dtype = np.float32 if out.dtype.kind == "f" else np.int32
    return Tensor(out.astype(dtype))


def meshgrid(*args, indexing="xy"):
    arrs = [a._v if isinstance(a, Tensor) else np.asarray(a) for a in args]
    return [Tensor(g) for g in np.meshgrid(*arrs, indexing=indexing)]


def _binary(fn):
    def op(a, b, name=None):
        x, y = _pair(a, b)
        with np.errstate(all="ignore"):
            return Tensor(fn(x, y))
    return op


add = _binary(np.add)
subtract = _binary(np.subtract)
multiply = _binary(np.multiply)
divide = _binary(np.true_divide)
maximum = _binary(np.maximum)
minimum = _binary(np.minimum)
less = _binary(np.less)
greater = _binary(np.greater)
pow = _binary(np.power)  # noqa: A001


def _unary(fn):
    def op(x, name=None):
        with np.errstate(all="ignore"):
            return Tensor(fn(_t(x)._v))
    return op


square = _unary(np.square)
abs = _unary(np.abs)  # noqa: A001
exp = _unary(np.exp)
sqrt = _unary(np.sqrt)
zeros_like = _unary(np.zeros_like)


def where(cond, x, y):
    c = cond._v if isinstance(cond, Tensor) else np.asarray(cond)
    a, b = _pair(x, y)
    return Tensor(np.where(c, a, b))


def _reduce(fn):
    def op(x, axis=None, keepdims=False):
        v = _t(x)._v
        return Tensor(fn(v, axis=axis, keepdims=keepdims).astype(v.dtype))
    return op


reduce_sum = _reduce(np.sum)
reduce_max = _reduce(np.max)
reduce_min = _reduce(np.min)
reduce_mean = _reduce(np.mean)


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_t(x)._v, axis))


def squeeze(x, axis=None):
    return Tensor(np.squeeze(_t(x)._v, axis))


def _seq(values):
    ts = [_t(v) for v in values]
    dt = ts[0].dtype
    for t in ts:
        if t.dtype != dt:
            raise TypeError("stub tf: mixed dtypes in stack/concat")
    return [t._v for t in ts]


def stack(values, axis=0):
    return Tensor(np.stack(_seq(values), axis=axis))


def concat(values, axis):
    return Tensor(np.concatenate(_seq(values), axis=axis))


def clip_by_value(x, lo, hi):
    v = _t(x)._v
    return Tensor(np.clip(v, np.asarray(lo).astype(v.dtype), np.asarray(hi).astype(v.dtype)))


class _Math:
    sqrt = staticmethod(sqrt)
    exp = staticmethod(exp)
    add = staticmethod(add)
    multiply = staticmethod(multiply)
    maximum = staticmethod(maximum)
    minimum = staticmethod(minimum)
    square = staticmethod(square)
    abs = staticmethod(abs)
    pow = staticmethod(pow)
    reduce_sum = staticmethod(reduce_sum)
    reduce_max = staticmethod(reduce_max)

    @staticmethod
    def log(x):
        with np.errstate(all="ignore"):
            return Tensor(np.log(_t(x)._v))

    @staticmethod
    def argmax(x, axis=None, output_type=np.int64):
        return Tensor(np.argmax(_t(x)._v, axis=axis).astype(output_type))

    @staticmethod
    def divide_no_nan(a, b):
        x, y = _pair(a, b)
        with np.errstate(all="ignore"):
            return Tensor(np.where(y == 0, np.zeros_like(x), x / y))


math = _Math()
argmax = _Math.argmax


class _NN:
    @staticmethod
    def sigmoid(x):
        v = _t(x)._v
        with np.errstate(all="ignore"):
            one = np.asarray(1, dtype=v.dtype)
            return Tensor(one / (one + np.exp(-v)))

    @staticmethod
    def relu(x):
        v = _t(x)._v
        return Tensor(np.maximum(v, np.asarray(0, dtype=v.dtype)))


nn = _NN()


class _Random:
    """tf.random.uniform(()) draws from np.random (seed it with np.random.seed for reproducible reference runs)."""
    @staticmethod
    def uniform(shape=(), minval=0.0, maxval=1.0, dtype=None):
        return Tensor(np.random.uniform(minval, maxval, size=tuple(shape)).astype(np.float32))


class _Image:
    @staticmethod
    def flip_left_right(image):
        return Tensor(np.asarray(_t(image)._v)[..., :, ::-1, :])


random = _Random()
image = _Image()


def constant_initializer(value=0):
    return ("constant_initializer", value)


def device(name):
    class _Ctx:
        def __enter__(self): return self
        def __exit__(self, *a): return False
    return _Ctx()
