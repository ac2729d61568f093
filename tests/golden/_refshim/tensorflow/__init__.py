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
  * tf.GradientTape (used by tests/golden/make_golden_grad.py only): reverse-mode autodiff over
    the ops above with TensorFlow's rules -- tf.where passes the gradient to the selected branch,
    tf.clip_by_value passes it where lo <= x <= hi, tf.argmin / comparisons / integer gathers carry
    none, tf.stop_gradient cuts it.  Forward values stay fp32; the backward pass accumulates in
    float64 on those values (the gradient every fp32 implementation approximates).
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

    def __init__(self, a, node=None):
        self._a = np.asarray(a)
        self._node = node            # autodiff graph node (None: constant w.r.t. the watched tensors)

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
        shp = self._a.shape

        def vjp(g):
            z = np.zeros(shp, np.float64)
            z[item] = g
            return z
        return _mk(self._a[item], [(self, vjp)])

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

    def __add__(self, o): return _binary('add', self, o, False)
    def __radd__(self, o): return _binary('add', self, o, True)
    def __sub__(self, o): return _binary('sub', self, o, False)
    def __rsub__(self, o): return _binary('sub', self, o, True)
    def __mul__(self, o): return _binary('mul', self, o, False)
    def __rmul__(self, o): return _binary('mul', self, o, True)
    def __truediv__(self, o): return _binary('div', self, o, False)
    def __rtruediv__(self, o): return _binary('div', self, o, True)
    def __neg__(self): return _mk(-self._a, [(self, lambda g: -g)])
    def __lt__(self, o): return Tensor(self._a < self._coerce(o))
    def __le__(self, o): return Tensor(self._a <= self._coerce(o))
    def __gt__(self, o): return Tensor(self._a > self._coerce(o))
    def __ge__(self, o): return Tensor(self._a >= self._coerce(o))
    def __eq__(self, o): return Tensor(self._a == self._coerce(o))
    def __ne__(self, o): return Tensor(self._a != self._coerce(o))
    __hash__ = None



# ------------------------------------------------------------------------------------------
# reverse-mode autodiff (tf.GradientTape stand-in)
# ------------------------------------------------------------------------------------------
_TAPES = []


class _Node(object):
    __slots__ = ('parents',)

    def __init__(self, parents):
        self.parents = parents           # [(parent Tensor, vjp: upstream grad -> grad w.r.t. parent)]


def _mk(out, parents):
    """Result tensor of an op; recorded on the tape if a tape is active and an operand is tracked."""
    if _TAPES:
        live = [(t, f) for t, f in parents if isinstance(t, Tensor) and t._node is not None]
        if live:
            return Tensor(out, _Node(live))
    return Tensor(out)


def _unbroadcast(g, shape):
    g = np.asarray(g, np.float64)
    while g.ndim > len(shape):
        g = g.sum(0)
    for i, n in enumerate(shape):
        if n == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdims=True)
    return g


def _binary(kind, self, other, reflected):
    b = self._coerce(other)                      # ndarray of the tensor's dtype, or a scalar of that dtype
    a = self._a
    x, y = (b, a) if reflected else (a, b)       # x (op) y
    if kind == 'add':
        out = x + y
    elif kind == 'sub':
        out = x - y
    elif kind == 'mul':
        out = x * y
    else:
        out = x / y
    x64, y64 = np.asarray(x, np.float64), np.asarray(y, np.float64)

    def vjp_x(g):
        if kind in ('add', 'sub'):
            return g
        return g * y64 if kind == 'mul' else g / y64

    def vjp_y(g):
        if kind == 'add':
            return g
        if kind == 'sub':
            return -g
        return g * x64 if kind == 'mul' else -g * x64 / (y64 * y64)

    parents = []
    first, second = (other, self) if reflected else (self, other)
    for t, f, v in ((first, vjp_x, x), (second, vjp_y, y)):
        if isinstance(t, Tensor):
            shp = np.shape(v)
            parents.append((t, (lambda f_, shp_: lambda g: _unbroadcast(f_(g), shp_))(f, shp)))
    return _mk(out, parents)


class GradientTape(object):
    """`with tf.GradientTape() as tape: tape.watch(x); y = f(x)` then `tape.gradient(y, [x], output_gradients)`."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        pass

    def __enter__(self):
        _TAPES.append(self)
        return self

    def __exit__(self, *exc):
        _TAPES.remove(self)
        return False

    def watch(self, t):
        for x in (t if isinstance(t, (list, tuple)) else [t]):
            if x._node is None:
                x._node = _Node([])

    def gradient(self, target, sources, output_gradients=None):
        targets = list(target) if isinstance(target, (list, tuple)) else [target]
        ups = list(output_gradients) if isinstance(output_gradients, (list, tuple)) else [output_gradients] * len(targets)
        grads, order, seen = {}, [], set()

        def visit(t):
            if t._node is None or id(t._node) in seen:
                return
            seen.add(id(t._node))
            for p, _ in t._node.parents:
                visit(p)
            order.append(t)
        import sys
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
        for t, u in zip(targets, ups):
            if t._node is None:
                continue
            visit(t)
            g = np.ones(t._a.shape, np.float64) if u is None else np.asarray(_raw(u), np.float64)
            grads[id(t._node)] = grads.get(id(t._node), 0) + g
        for t in reversed(order):
            g = grads.get(id(t._node))
            if g is None:
                continue
            for p, f in t._node.parents:
                grads[id(p._node)] = grads.get(id(p._node), 0) + f(g)
        single = not isinstance(sources, (list, tuple))
        res = []
        for s_ in ([sources] if single else sources):
            g = grads.get(id(s_._node)) if s_._node is not None else None
            res.append(None if g is None else Tensor(np.asarray(g, np.float64)))
        return res[0] if single else res

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
    a = _t(x)
    if np.issubdtype(np.dtype(dtype), np.floating) and np.issubdtype(a.dtype, np.floating):
        return _mk(a.astype(dtype), [(x, lambda g: g)])
    return Tensor(a.astype(dtype))


def zeros(shape, dtype=np.float32):
    return Tensor(np.zeros(shape, dtype=dtype))


def zeros_like(x):
    return Tensor(np.zeros_like(_t(x)))


def ones_like(x):
    return Tensor(np.ones_like(_t(x)))


def square(x):
    a = _t(x)
    a64 = a.astype(np.float64)
    return _mk(a * a, [(x, lambda g: 2.0 * a64 * g)])


def sqrt(x):
    r = np.sqrt(_t(x))
    r64 = r.astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        return _mk(r, [(x, lambda g: g / (2.0 * r64))])


def _via_f64(fn, x, dfn):
    a = _t(x)
    a64 = a.astype(np.float64)
    return _mk(fn(a64).astype(a.dtype), [(x, lambda g: g * dfn(a64))])


def cos(x): return _via_f64(np.cos, x, lambda a: -np.sin(a))
def sin(x): return _via_f64(np.sin, x, np.cos)
def atan(x): return _via_f64(np.arctan, x, lambda a: 1.0 / (1.0 + a * a))


def where(cond, x, y):
    c = _t(cond)
    xr, yr = _raw(x), _raw(y)
    xa = xr if (isinstance(xr, np.ndarray) and xr.ndim > 0) else None
    ya = yr if (isinstance(yr, np.ndarray) and yr.ndim > 0) else None
    like = xa if xa is not None else ya
    xa = _t(x, like)
    ya = _t(y, like)
    out = np.where(c, xa, ya)
    cb = np.broadcast_to(c, out.shape)
    # TF: the gradient flows to the selected branch only
    return _mk(out, [(x, lambda g: _unbroadcast(np.where(cb, g, 0.0), np.shape(xa))),
                     (y, lambda g: _unbroadcast(np.where(cb, 0.0, g), np.shape(ya)))])


def logical_and(a, b):
    return Tensor(np.logical_and(_t(a), _t(b)))


def stack(values, axis=0):
    arrs = [_t(v) for v in values]
    parents = [(v, (lambda i: lambda g: np.take(g, i, axis=axis))(i)) for i, v in enumerate(values)]
    return _mk(np.stack(arrs, axis=axis), parents)


def concat(values, axis):
    arrs = [_t(v) for v in values]
    offs = np.cumsum([0] + [a.shape[axis] for a in arrs])
    parents = [(v, (lambda i: lambda g: np.take(g, np.arange(offs[i], offs[i + 1]), axis=axis))(i))
               for i, v in enumerate(values)]
    return _mk(np.concatenate(arrs, axis=axis), parents)


def tile(x, multiples):
    a = _t(x)
    m = tuple(int(k) for k in _t(multiples))

    def vjp(g):
        shp = []
        for k, n in zip(m, a.shape):
            shp += [k, n]
        return g.reshape(shp).sum(axis=tuple(range(0, 2 * len(m), 2)))
    return _mk(np.tile(a, m), [(x, vjp)])


def reshape(x, shape):
    a = _t(x)
    return _mk(np.reshape(a, shape), [(x, lambda g: g.reshape(a.shape))])


def expand_dims(x, axis):
    a = _t(x)
    return _mk(np.expand_dims(a, axis), [(x, lambda g: g.reshape(a.shape))])


def argmin(x, axis):
    return Tensor(np.argmin(_t(x), axis=axis).astype(np.int64))


def gather(params, indices):
    return Tensor(_t(params)[_t(indices)])


def clip_by_value(x, lo, hi):
    a = _t(x)
    lo_, hi_ = a.dtype.type(lo), a.dtype.type(hi)
    inside = (a >= lo_) & (a <= hi_)                 # TF: gradient passes on the closed interval
    return _mk(np.minimum(np.maximum(a, lo_), hi_), [(x, lambda g: np.where(inside, g, 0.0))])


def stop_gradient(x):
    return Tensor(_t(x)) if isinstance(x, Tensor) else x


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
