"""NumPy-backed stand-in for the handful of TensorFlow ops the reference hot path uses.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installable in this image (no
network), so the UNMODIFIED reference source (/root/reference/dynamics_and_models.py)
is executed on top of this module by tests/golden/make_golden.py to produce the
golden vectors under tests/golden/.  Nothing in the product imports this.

Semantics reproduced (what the reference relies on):
  * every tensor op is fp32, one IEEE rounding per op, no FMA contraction
    (NumPy element-wise ops on float32 arrays behave exactly so);
  * a Python / NumPy scalar combined with an fp32 tensor is first converted to
    fp32 (TF's convert_to_tensor with the tensor's dtype);
  * tf.argmin returns the FIRST minimum (np.argmin does too), as int64;
  * tf.cos / tf.sin / tf.atan: TF's Eigen kernels are ~1 ulp approximations that
    cannot be reproduced bit-for-bit offline; the shim evaluates them in float64
    and rounds to fp32 -- the value every fp32 implementation approximates.
    This is the same policy oracle/ follows, so oracle-vs-golden is bit-exact.
"""
import contextlib

import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64


def _raw(x):
    return x._a if isinstance(x, Tensor) else x


class Tensor(object):
    """Minimal EagerTensor look-alike over a NumPy array."""
    __array_priority__ = 1000  # make ndarray defer to our reflected operators
    __array_ufunc__ = None

    def __init__(self, a):
        self._a = np.asarray(a)

    # -- introspection ------------------------------------------------------
    def numpy(self):
        return self._a

    @property
    def shape(self):
        return self._a.shape

    @property
    def dtype(self):
        return self._a.dtype

    def __len__(self):
        return len(self._a)

    def __getitem__(self, item):
        return Tensor(self._a[item])

    def __iter__(self):
        for i in range(len(self._a)):
            yield Tensor(self._a[i])

    def __repr__(self):
        return 'shim.Tensor(%r)' % (self._a,)

    def __bool__(self):
        return bool(self._a)

    def __float__(self):
        return float(self._a)

    def __int__(self):
        return int(self._a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    # -- arithmetic with TF dtype rules ------------------------------------
    def _coerce(self, other):
        o = _raw(other)
        if isinstance(o, np.ndarray) and o.ndim > 0:
            if o.dtype != self._a.dtype:
                # TF would raise on mixed dtypes; the reference only mixes
                # float32 arrays, so be strict to catch shim mis-use.
                if np.issubdtype(self._a.dtype, np.floating) and np.issubdtype(o.dtype, np.floating):
                    raise TypeError('dtype mismatch %s vs %s' % (self._a.dtype, o.dtype))
                o = o.astype(self._a.dtype)
            return o
        # Python / NumPy scalar -> tensor's dtype (one rounding, like TF)
        return self._a.dtype.type(o)

    def __add__(self, o): return Tensor(self._a + self._coerce(o))
    def __radd__(self, o): return Tensor(self._coerce(o) + self._a)
    def __sub__(self, o): return Tensor(self._a - self._coerce(o))
    def __rsub__(self, o): return Tensor(self._coerce(o) - self._a)
    def __mul__(self, o): return Tensor(self._a * self._coerce(o))
    def __rmul__(self, o): return Tensor(self._coerce(o) * self._a)
    def __truediv__(self, o): return Tensor(self._a / self._coerce(o))
    def __rtruediv__(self, o): return Tensor(self._coerce(o) / self._a)
    def __neg__(self): return Tensor(-self._a)
    def __lt__(self, o): return Tensor(self._a < self._coerce(o))
    def __le__(self, o): return Tensor(self._a <= self._coerce(o))
    def __gt__(self, o): return Tensor(self._a > self._coerce(o))
    def __ge__(self, o): return Tensor(self._a >= self._coerce(o))
    def __eq__(self, o): return Tensor(self._a == self._coerce(o))
    def __ne__(self, o): return Tensor(self._a != self._coerce(o))
    __hash__ = None


def _t(x, like=None):
    """To raw ndarray; scalars take `like`'s dtype."""
    r = _raw(x)
    if isinstance(r, np.ndarray) and r.ndim > 0:
        return r
    if like is not None:
        return np.asarray(r, dtype=like.dtype)
    return np.asarray(r)


def convert_to_tensor(value, dtype=None):
    a = np.asarray(_raw(value))
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(np.float32)  # TF default float is fp32
    return Tensor(a)


def constant(value, dtype=None):
    a = np.asarray(value)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(np.float32)
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return Tensor(a)


def cast(x, dtype):
    return Tensor(_t(x).astype(dtype))


def zeros(shape, dtype=np.float32):
    return Tensor(np.zeros(shape, dtype=dtype))


def zeros_like(x):
    return Tensor(np.zeros_like(_t(x)))


def ones_like(x):
    return Tensor(np.ones_like(_t(x)))


def square(x):
    a = _t(x)
    return Tensor(a * a)


def sqrt(x):
    return Tensor(np.sqrt(_t(x)))


def _via_f64(fn, x):
    a = _t(x)
    return Tensor(fn(a.astype(np.float64)).astype(a.dtype))


def cos(x): return _via_f64(np.cos, x)
def sin(x): return _via_f64(np.sin, x)
def atan(x): return _via_f64(np.arctan, x)


def where(cond, x, y):
    c = _t(cond)
    xr, yr = _raw(x), _raw(y)
    xa = xr if (isinstance(xr, np.ndarray) and xr.ndim > 0) else None
    ya = yr if (isinstance(yr, np.ndarray) and yr.ndim > 0) else None
    like = xa if xa is not None else ya
    xa = _t(x, like)
    ya = _t(y, like)
    return Tensor(np.where(c, xa, ya))


def logical_and(a, b):
    return Tensor(np.logical_and(_t(a), _t(b)))


def stack(values, axis=0):
    return Tensor(np.stack([_t(v) for v in values], axis=axis))


def concat(values, axis):
    return Tensor(np.concatenate([_t(v) for v in values], axis=axis))


def tile(x, multiples):
    return Tensor(np.tile(_t(x), tuple(int(m) for m in _t(multiples))))


def reshape(x, shape):
    return Tensor(np.reshape(_t(x), shape))


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_t(x), axis))


def argmin(x, axis):
    return Tensor(np.argmin(_t(x), axis=axis).astype(np.int64))


def gather(params, indices):
    return Tensor(_t(params)[_t(indices)])


def clip_by_value(x, lo, hi):
    a = _t(x)
    return Tensor(np.minimum(np.maximum(a, a.dtype.type(lo)), a.dtype.type(hi)))


def stop_gradient(x):
    return x


def shape(x):
    return np.asarray(_t(x).shape)


@contextlib.contextmanager
def name_scope(name):
    yield name


def function(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


class _Threading(object):
    @staticmethod
    def set_inter_op_parallelism_threads(n):
        pass

    @staticmethod
    def set_intra_op_parallelism_threads(n):
        pass


class _Config(object):
    threading = _Threading()


config = _Config()
