"""A NumPy stand-in for the TensorFlow-1.x graph API, just large enough to EXECUTE THE REFERENCE'S
OWN SOURCE FILES ``zhusuan/hmc.py`` and ``zhusuan/sgmcmc.py`` unmodified (TEST INFRASTRUCTURE ONLY).

Why it exists.  TensorFlow cannot be installed here (no wheel, no network), so round 1 could only
restate hmc.py / sgmcmc.py in NumPy (oracle/hmc.py, oracle/sgmcmc.py) -- "parity unpinned by the
reference".  With this module registered as ``tensorflow`` the reference files themselves run:
their control flow, variable updates, quirks (``mu = 10 * eps0``, the EWMV update order, the
step-size search loop, ...) are then the reference's, not a restatement.  oracle/tf_shim/
make_ref_golden.py drives them and writes tests/golden/ref_*.npz; tests/test_ref_pins.py checks the
oracle (and, on the GPU, the CUDA path) against those vectors.

What it is.  A lazily evaluated dataflow graph with TF-1.x semantics:
  * every op builds a ``Tensor`` holding a closure; ``Session.run(fetches, feed_dict)`` evaluates
    the fetches in order with per-run memoisation (an op executes at most once per run) after its
    control dependencies (``tf.control_dependencies``);
  * ``Variable`` reads return the value at the moment of the FIRST read in a run (the reference
    only relies on orders that its data / control dependencies enforce; every use of an updated
    value goes through the assign op's output);
  * ``tf.cond`` traces both branches at graph-construction time and executes only the taken one;
    ``tf.while_loop`` re-traces its body on concrete loop values at run time;
  * ``tf.gradients`` is reverse-mode differentiation over the recorded ops (the handful the
    test log-joints use);
  * ``tf.random_normal`` / ``tf.random_uniform`` pop arrays injected with ``set_noise`` (TF's own
    Philox streams are irrelevant to parity: all parity runs inject noise);
  * arithmetic is NumPy float32 (IEEE +,-,*,/ and sqrt are bit-identical to TF-CPU's; exp / pow /
    reductions may differ from Eigen's kernels in the last ulp).
Nothing outside hmc.py / sgmcmc.py's needs is implemented; an unknown attribute raises.
"""
import contextlib

import numpy as np

__version__ = "1.x-numpy-shim"

float32, float64, int32, int64, bool = np.float32, np.float64, np.int32, np.int64, np.bool_
float16, int16, int8, uint8 = np.float16, np.int16, np.int8, np.uint8
_SERIAL = [0]
_CTRL_STACK = [[]]


# ------------------------------------------------------------------------------------------------
class Dimension(int):
    """tf.Dimension: an int whose `.value` is itself (monte_carlo.py:173 reads `.value`)."""
    value = property(lambda self: int(self))


class TensorShape(object):
    def __init__(self, dims):
        self._dims = None if dims is None else [None if d is None else Dimension(d) for d in dims]

    @property
    def ndims(self):
        return None if self._dims is None else len(self._dims)

    def as_list(self):
        return list(self._dims)

    def is_fully_defined(self):
        return self._dims is not None and all(d is not None for d in self._dims)

    def is_compatible_with(self, other):
        o = other if isinstance(other, TensorShape) else TensorShape(other)
        if self._dims is None or o._dims is None:
            return True
        return len(self._dims) == len(o._dims) and all(
            a is None or b is None or a == b for a, b in zip(self._dims, o._dims))

    def concatenate(self, other):
        return TensorShape(self._dims + TensorShape(other)._dims if not isinstance(
            other, TensorShape) else self._dims + other._dims)

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims)

    def __getitem__(self, k):
        return TensorShape(self._dims[k]) if isinstance(k, slice) else self._dims[k]

    def __bool__(self):                      # TF 1.x: unknown rank is falsy, any known shape truthy
        return self._dims is not None

    __nonzero__ = __bool__

    def __eq__(self, other):
        return list(self._dims) == list(TensorShape(other)._dims if not isinstance(
            other, TensorShape) else other._dims)

    def __repr__(self):
        return "TensorShape(%r)" % (self._dims,)


class _Ctx(object):
    """One Session.run: memo of evaluated tensors, the feed, variable snapshots."""

    def __init__(self, feed, parent=None, floor=None, peek=False):
        self.feed = feed
        self.peek = peek            # static-shape inference, see _is_peek()
        self.memo = {}
        self.parent = parent
        self.floor = floor          # tensors created before `floor` belong to the parent

    def _home(self, t):
        c = self
        while c.parent is not None and t.serial < c.floor:
            c = c.parent
        return c

    def eval(self, t):
        t = convert_to_tensor(t)
        home = self._home(t)
        if id(t) in home.memo:
            return home.memo[id(t)][1]
        for d in t.deps:
            self.eval(d)
        v = t.fn(self)
        home.memo[id(t)] = (t, v)
        return v


def _is_peek(c):
    """Static-shape inference (`Tensor.get_shape`) evaluates a tensor in a scratch context.  There
    the graph must be free of side effects and cheap: random ops yield zeros without consuming
    injected noise, assigns do not write, unfed placeholders read as zeros, `cond` takes its first
    branch and `while_loop` returns its initial loop variables (both shape-preserving in TF)."""
    while c is not None:
        if c.peek:
            return True
        c = c.parent
    return False


class Tensor(object):
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), op="op", vjp=None, dtype=None, name=None):
        self.fn = fn
        self.inputs = tuple(inputs)
        self.op = op
        self.vjp = vjp              # (cotangent Tensor) -> list of cotangent Tensors / None per input
        self._dtype = dtype
        self.name = name
        _SERIAL[0] += 1
        self.serial = _SERIAL[0]
        self.deps = tuple(_CTRL_STACK[-1])

    # --- static information (computed by a throw-away evaluation: only used on pure tensors) ---
    def _peek(self):
        return _Ctx({}, peek=True).eval(self)

    def get_shape(self):
        return TensorShape(np.shape(self._peek()))

    shape = property(get_shape)

    def set_shape(self, shape):                 # static hints carry no information here
        pass

    @property
    def dtype(self):
        return self._dtype if self._dtype is not None else np.float32

    def eval(self, feed_dict=None, session=None):
        return Session().run(self, feed_dict)

    # --- operators ---
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __neg__(self): return negative(self)
    def __pow__(self, o): return pow(self, o)
    def __rpow__(self, o): return pow(o, self)
    def __abs__(self): return abs(self)
    def __lt__(self, o): return less(self, o)
    def __gt__(self, o): return greater(self, o)
    def __le__(self, o): return _cmp(np.less_equal, self, o)
    def __ge__(self, o): return _cmp(np.greater_equal, self, o)
    def __getitem__(self, k): return _unary(lambda a: a[k], self, "getitem")
    __hash__ = object.__hash__

    def __bool__(self):
        raise TypeError("a graph Tensor has no truth value (use tf.cond)")

    __nonzero__ = __bool__


def _like(value, ref):
    """Python scalars take the dtype of the tensor they meet (TF constant conversion)."""
    if isinstance(value, Tensor):
        return value
    if isinstance(ref, Tensor) and isinstance(value, (int, float, np.floating, np.integer)) \
            and not isinstance(value, (np.bool_, type(True))):
        dt = ref._dtype               # (never evaluated: dtypes are inferred structurally)
        if dt is not None and dt is not np.bool_:
            return constant(value, dtype=dt)
    return convert_to_tensor(value)


def _dt(*ts):
    """Result dtype of an op from its inputs' declared dtypes (None when unknown)."""
    ds = [t._dtype for t in ts if isinstance(t, Tensor) and t._dtype is not None]
    if not ds:
        return None
    return np.result_type(*ds).type


def _f32(a):
    a = np.asarray(a)
    return a.astype(np.float32) if a.dtype == np.float64 else a


def convert_to_tensor(value, dtype=None, name=None, preferred_dtype=None):
    for typ, conv in _TENSOR_CONVERSIONS:          # Tensor-likes registered by the reference
        if isinstance(value, typ):
            # dtype objects here are NumPy types (no .is_compatible_with): convert, then cast
            return convert_to_tensor(conv(value, dtype=None, name=name, as_ref=False), dtype)
    if isinstance(value, Tensor):
        if dtype is not None and value._dtype is not None and value._dtype is not dtype:
            return cast(value, dtype)
        return value
    if isinstance(value, TensorShape):
        value = value.as_list()
    if isinstance(value, (list, tuple)) and any(isinstance(v, Tensor) for v in value):
        parts = [convert_to_tensor(v) for v in value]            # tf.stack (auto-packing)
        out = Tensor(lambda c: np.stack([np.asarray(c.eval(p)) for p in parts]),
                     inputs=tuple(parts), op="pack", dtype=_dt(*parts))
        return cast(out, dtype) if dtype is not None and out._dtype is not dtype else out
    return constant(value, dtype=dtype, name=name)


def constant(value, dtype=None, shape=None, name=None):
    if dtype is None:
        a = np.asarray(value)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    else:
        a = np.asarray(value, dtype=dtype)
    if shape is not None:
        a = np.broadcast_to(a, tuple(shape)).copy()
    return Tensor(lambda c, a=a: a, op="const", dtype=a.dtype.type, name=name)


def placeholder(dtype, shape=None, name=None):
    t = Tensor(None, op="placeholder", dtype=dtype, name=name)

    def fn(c, t=t):
        cc = c
        while cc is not None:
            if t in cc.feed:
                return np.asarray(cc.feed[t], dtype=dtype)
            if cc.peek and cc.parent is None:
                return np.zeros(tuple(shape) if shape is not None else (), dtype=dtype)
            cc = cc.parent
        raise ValueError("placeholder %r was not fed" % (name,))
    t.fn = fn
    return t


# ------------------------------------------------------------------------------------------------
class Variable(Tensor):
    def __init__(self, initial_value, name=None, trainable=True, dtype=None):
        if isinstance(initial_value, Tensor):
            # a real evaluation (initialisers draw from the injected noise), not a shape peek
            v = np.array(_Ctx({}).eval(initial_value))
        else:
            v = np.array(initial_value)
        if dtype is not None:
            v = v.astype(dtype)
        elif v.dtype == np.float64:
            v = v.astype(np.float32)
        elif v.dtype == np.int64:
            v = v.astype(np.int32)
        self.value = v
        Tensor.__init__(self, lambda c: self.value, op="variable", dtype=v.dtype.type, name=name)
        self.deps = ()

    def assign(self, value, use_locking=None):
        return assign(self, value)

    def assign_add(self, delta, use_locking=None):
        return assign(self, self + _like(delta, self))

    def load(self, value, session=None):
        self.value = np.array(value, dtype=self.value.dtype)

    def _peek(self):
        return self.value

    @property
    def initializer(self):
        return no_op()


def assign(ref, value, validate_shape=None, use_locking=None):
    value = _like(value, ref)

    def fn(c):
        v = np.array(c.eval(value), dtype=ref.value.dtype)
        if _is_peek(c):
            return v.reshape(ref.value.shape) if v.shape != ref.value.shape else v
        c.eval(ref)                        # pin the pre-assignment snapshot for this run
        ref.value = v.reshape(ref.value.shape) if v.shape != ref.value.shape else v
        return ref.value
    return Tensor(fn, inputs=(value,), op="assign", dtype=ref._dtype)


def global_variables_initializer():
    return no_op()


def no_op(name=None):
    return Tensor(lambda c: None, op="no_op")


def group(*ops, **kw):
    ops = [o for o in ops if o is not None]

    def fn(c):
        for o in ops:
            c.eval(o)
        return None
    return Tensor(fn, inputs=tuple(ops), op="group")


@contextlib.contextmanager
def control_dependencies(deps):
    deps = [d for d in (deps or []) if isinstance(d, Tensor)]
    _CTRL_STACK.append(_CTRL_STACK[-1] + deps)
    try:
        yield
    finally:
        _CTRL_STACK.pop()


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name


variable_scope = name_scope


# ------------------------------------------------------------------------------------------------
def _unbroadcast(g, like):
    """Sum the cotangent `g` (Tensor) back to the shape of input `like` (Tensor)."""
    def fn(c):
        gv, shape = np.asarray(c.eval(g)), np.shape(c.eval(like))
        while gv.ndim > len(shape):
            gv = gv.sum(0)
        for ax, n in enumerate(shape):
            if n == 1 and gv.shape[ax] != 1:
                gv = gv.sum(ax, keepdims=True)
        return gv.astype(np.asarray(c.eval(like)).dtype, copy=False)
    return Tensor(fn, inputs=(g, like), op="unbroadcast", dtype=like._dtype)


def _binary(npf, a, b, op, vjp=None):
    a, b = _like(a, b), _like(b, a)
    return Tensor(lambda c: npf(c.eval(a), c.eval(b)), inputs=(a, b), op=op, vjp=vjp,
                  dtype=_dt(a, b))


def _unary(npf, a, op, vjp=None, dtype=None):
    a = convert_to_tensor(a)
    return Tensor(lambda c: npf(c.eval(a)), inputs=(a,), op=op, vjp=vjp,
                  dtype=dtype if dtype is not None else a._dtype)


def add(a, b, name=None):
    a, b = _like(a, b), _like(b, a)
    return _binary(np.add, a, b, "add", lambda g: [_unbroadcast(g, a), _unbroadcast(g, b)])


def subtract(a, b, name=None):
    a, b = _like(a, b), _like(b, a)
    return _binary(np.subtract, a, b, "sub",
                   lambda g: [_unbroadcast(g, a), _unbroadcast(negative(g), b)])


def multiply(a, b, name=None):
    a, b = _like(a, b), _like(b, a)
    return _binary(np.multiply, a, b, "mul",
                   lambda g: [_unbroadcast(g * b, a), _unbroadcast(g * a, b)])


def divide(a, b, name=None):
    a, b = _like(a, b), _like(b, a)
    return _binary(np.true_divide, a, b, "div",
                   lambda g: [_unbroadcast(g / b, a), _unbroadcast(negative(g * a / (b * b)), b)])


div = truediv = realdiv = divide


def negative(a, name=None):
    return _unary(np.negative, a, "neg", lambda g: [negative(g)])


def square(a, name=None):
    a = convert_to_tensor(a)
    return _unary(np.square, a, "square", lambda g: [g * (a * 2.0)])


def sqrt(a, name=None):
    a = convert_to_tensor(a)
    out = _unary(np.sqrt, a, "sqrt")
    out.vjp = lambda g: [g * 0.5 / out]
    return out


def exp(a, name=None):
    a = convert_to_tensor(a)
    out = _unary(lambda x: np.exp(x).astype(np.asarray(x).dtype, copy=False), a, "exp")
    out.vjp = lambda g: [g * out]
    return out


def log(a, name=None):
    a = convert_to_tensor(a)
    return _unary(np.log, a, "log", lambda g: [g / a])


def abs(a, name=None):                                         # noqa: A001
    a = convert_to_tensor(a)
    return _unary(np.abs, a, "abs", lambda g: [g * sign(a)])


def sign(a, name=None):
    return _unary(np.sign, a, "sign")


def pow(a, b, name=None):                                      # noqa: A001
    a, b = _like(a, b), _like(b, a)
    return _binary(lambda x, y: np.power(x, y).astype(np.result_type(x, y), copy=False), a, b, "pow")


def mod(a, b, name=None):
    return _binary(np.mod, a, b, "mod")


floormod = mod


def minimum(a, b, name=None):
    a, b = _like(a, b), _like(b, a)
    return _binary(np.minimum, a, b, "minimum")


def maximum(a, b, name=None):
    a, b = _like(a, b), _like(b, a)
    return _binary(np.maximum, a, b, "maximum")


def _cmp(npf, a, b):
    a, b = _like(a, b), _like(b, a)
    return Tensor(lambda c: npf(c.eval(a), c.eval(b)), inputs=(a, b), op="cmp", dtype=np.bool_)


def less(a, b, name=None): return _cmp(np.less, a, b)
def greater(a, b, name=None): return _cmp(np.greater, a, b)
def equal(a, b, name=None): return _cmp(np.equal, a, b)
def logical_and(a, b, name=None): return _cmp(np.logical_and, a, b)
def logical_or(a, b, name=None): return _cmp(np.logical_or, a, b)
def logical_xor(a, b, name=None): return _cmp(np.logical_xor, a, b)
def logical_not(a, name=None): return _unary(np.logical_not, a, "not", dtype=np.bool_)
def is_finite(a, name=None): return _unary(np.isfinite, a, "is_finite", dtype=np.bool_)


def identity(a, name=None):
    return _unary(lambda x: x, a, "identity", lambda g: [g])


def stop_gradient(a, name=None):
    return _unary(lambda x: x, a, "stop_gradient", lambda g: [None])


def cast(a, dtype, name=None):
    a = convert_to_tensor(a)
    flt = np.dtype(dtype).kind == "f" and (a._dtype is None or np.dtype(a._dtype).kind == "f")
    return _unary(lambda x: np.asarray(x).astype(dtype), a, "cast",
                  (lambda g: [cast(g, a._dtype or np.float32)]) if flt else None, dtype=dtype)


to_float = lambda a, name=None: cast(a, np.float32)            # noqa: E731
to_int32 = lambda a, name=None: cast(a, np.int32)              # noqa: E731


def check_numerics(a, message, name=None):
    a = convert_to_tensor(a)

    def f(x):
        if not np.all(np.isfinite(x)):
            raise errors.InvalidArgumentError(None, None, message)
        return x
    return _unary(f, a, "check_numerics", lambda g: [g])


def _shape_arg(c, shape):
    if isinstance(shape, Tensor):
        return tuple(int(s) for s in np.atleast_1d(c.eval(shape)))
    if isinstance(shape, TensorShape):
        return tuple(shape.as_list())
    return tuple(int(c.eval(s)) if isinstance(s, Tensor) else int(s) for s in shape)


def zeros(shape, dtype=np.float32, name=None):
    return Tensor(lambda c: np.zeros(_shape_arg(c, shape), dtype), op="zeros", dtype=dtype)


def ones(shape, dtype=np.float32, name=None):
    return Tensor(lambda c: np.ones(_shape_arg(c, shape), dtype), op="ones", dtype=dtype)


def zeros_like(a, dtype=None, name=None):
    a = convert_to_tensor(a)
    return _unary(lambda x: np.zeros(np.shape(x), dtype or np.asarray(x).dtype), a, "zeros_like",
                  dtype=dtype)


def ones_like(a, dtype=None, name=None):
    a = convert_to_tensor(a)
    return _unary(lambda x: np.ones(np.shape(x), dtype or np.asarray(x).dtype), a, "ones_like",
                  dtype=dtype)


def shape(a, name=None, out_type=np.int32):                    # noqa: F811
    a = convert_to_tensor(a)
    return _unary(lambda x: np.asarray(np.shape(x), np.int32), a, "shape", dtype=np.int32)


def range(*args, **kw):                                        # noqa: A001
    ts = [convert_to_tensor(a) for a in args]
    return Tensor(lambda c: np.arange(*[int(c.eval(t)) for t in ts], dtype=np.int32),
                  inputs=tuple(ts), op="range", dtype=np.int32)


def expand_dims(a, axis=None, name=None, dim=None):
    a = convert_to_tensor(a)
    ax = axis if axis is not None else dim
    if isinstance(ax, Tensor):
        return Tensor(lambda c: np.expand_dims(c.eval(a), int(c.eval(ax))), inputs=(a, ax),
                      op="expand_dims", vjp=lambda g: [reshape(g, shape(a)), None],
                      dtype=a._dtype)
    return _unary(lambda x: np.expand_dims(x, ax), a, "expand_dims",
                  lambda g: [reshape(g, shape(a))])


def reshape(a, shp, name=None):
    a = convert_to_tensor(a)
    return Tensor(lambda c: np.reshape(c.eval(a), _shape_arg(c, shp)), inputs=(a,), op="reshape",
                  vjp=lambda g: [reshape(g, shape(a))], dtype=a._dtype)


def tile(a, multiples, name=None):
    a = convert_to_tensor(a)
    return Tensor(lambda c: np.tile(c.eval(a), _shape_arg(c, multiples)), inputs=(a,), op="tile",
                  dtype=a._dtype)


def where(cond, x=None, y=None, name=None):
    cond, x, y = convert_to_tensor(cond), convert_to_tensor(x), convert_to_tensor(y)

    def fn(c):
        cv, xv, yv = c.eval(cond), c.eval(x), c.eval(y)
        if np.ndim(cv) == 1 and np.ndim(xv) > 1:         # TF: a vector condition selects rows
            cv = cv.reshape((-1,) + (1,) * (np.ndim(xv) - 1))
        return np.where(cv, xv, yv)
    return Tensor(fn, inputs=(cond, x, y), op="where", dtype=_dt(x, y))


def _axes(c, axis, ndim):
    if axis is None:
        return None
    if isinstance(axis, Tensor):
        axis = c.eval(axis)
    ax = tuple(int(a) for a in np.atleast_1d(axis))
    return tuple(a % ndim for a in ax) if ndim else ax


def _reduce(npf, a, axis, keepdims, op, vjp_scale):
    a = convert_to_tensor(a)

    def fn(c):
        x = np.asarray(c.eval(a))
        ax = _axes(c, axis, x.ndim)
        if ax is not None and len(ax) == 0:
            return x
        kd = True if keepdims else False
        if npf in (np.max, np.min):
            return npf(x, axis=ax, keepdims=kd)
        return npf(x, axis=ax, keepdims=kd, dtype=x.dtype if x.dtype.kind == "f" else None)
    out = Tensor(fn, inputs=(a,), op=op, dtype=a._dtype)

    def vjp(g):
        def gfn(c):
            x = np.asarray(c.eval(a))
            gv = np.asarray(c.eval(g))
            ax = _axes(c, axis, x.ndim)
            n = 1.0
            if ax is None:
                ax = tuple(np.arange(x.ndim))
            if not keepdims:
                for d in sorted(ax):
                    gv = np.expand_dims(gv, d)
            for d in ax:
                n *= x.shape[d]
            gv = np.broadcast_to(gv, x.shape)
            return (gv / x.dtype.type(n) if vjp_scale else gv).astype(x.dtype, copy=False)
        return [Tensor(gfn, inputs=(g, a), op=op + "_grad", dtype=a._dtype)]
    out.vjp = vjp
    return out


def reduce_sum(a, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    axis = axis if axis is not None else reduction_indices
    return _reduce(np.sum, a, axis, keepdims or keep_dims, "reduce_sum", False)


def reduce_mean(a, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    axis = axis if axis is not None else reduction_indices
    return _reduce(np.mean, a, axis, keepdims or keep_dims, "reduce_mean", True)


def add_n(inputs, name=None):
    out = inputs[0]
    for t in inputs[1:]:
        out = add(out, t)
    return out


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = convert_to_tensor(a), convert_to_tensor(b)
    ta = (lambda x: np.swapaxes(x, -1, -2)) if transpose_a else (lambda x: x)
    tb = (lambda x: np.swapaxes(x, -1, -2)) if transpose_b else (lambda x: x)
    out = Tensor(lambda c: np.matmul(ta(c.eval(a)), tb(c.eval(b))), inputs=(a, b), op="matmul",
                 dtype=_dt(a, b))
    if not transpose_a and not transpose_b:
        out.vjp = lambda g: [matmul(g, b, transpose_b=True), matmul(a, g, transpose_a=True)]
    return out


# ------------------------------------------------------------------------------------------------
def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None, strict=False):
    """Both branches are traced now (graph construction); only the taken one ever executes, so
    side effects created inside a branch (assigns) are conditional, as in TF."""
    true_fn, false_fn = true_fn or fn1, false_fn or fn2
    pred = convert_to_tensor(pred)
    t_out, f_out = true_fn(), false_fn()
    is_list = isinstance(t_out, (list, tuple))
    t_list = [convert_to_tensor(x) for x in (t_out if is_list else [t_out])]
    f_list = [convert_to_tensor(x) for x in (f_out if is_list else [f_out])]
    outs = []
    for tt, ff in zip(t_list, f_list):
        def fn(c, tt=tt, ff=ff):
            if _is_peek(c):
                return c.eval(tt)
            return c.eval(tt) if builtins_bool(c.eval(pred)) else c.eval(ff)
        outs.append(Tensor(fn, inputs=(pred, tt, ff), op="cond", dtype=_dt(tt, ff)))
    return outs if is_list else outs[0]


def builtins_bool(x):
    return True if np.asarray(x).item() else False


def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10,   # noqa: F811
               back_prop=True, swap_memory=False, name=None, maximum_iterations=None):
    flat = []

    def flatten(v):
        if isinstance(v, (list, tuple)):
            return [flatten(x) for x in v]
        flat.append(convert_to_tensor(v))
        return len(flat) - 1
    structure = flatten(list(loop_vars))
    n = len(flat)

    def rebuild(struct, vals):
        return [rebuild(s, vals) if isinstance(s, list) else vals[s] for s in struct]

    def run(c):
        vals = [c.eval(t) for t in flat]
        it = 0
        while not _is_peek(c):
            floor = _SERIAL[0] + 1
            sub = _Ctx(c.feed, parent=c, floor=floor)     # this iteration's temporaries
            consts = [Tensor(lambda cc, v=v: v, op="loop_var",
                             dtype=np.asarray(v).dtype.type) for v in vals]
            args = rebuild(structure, consts)
            if not builtins_bool(sub.eval(convert_to_tensor(cond(*args)))):
                break
            if maximum_iterations is not None and it >= maximum_iterations:
                break
            out = body(*args)
            out_flat = []

            def fl(v):
                if isinstance(v, (list, tuple)):
                    for x in v:
                        fl(x)
                else:
                    out_flat.append(convert_to_tensor(v))
            fl(list(out))
            assert len(out_flat) == n, "while_loop: body changed the structure"
            vals = [sub.eval(t) for t in out_flat]
            it += 1
        return vals
    all_out = Tensor(run, inputs=tuple(flat), op="while_all")
    outs = [Tensor(lambda c, k=k: c.eval(all_out)[k], inputs=(all_out,), op="while",
                   dtype=flat[k]._dtype) for k in np.arange(n)]
    rebuilt = rebuild(structure, outs)
    return rebuilt if isinstance(loop_vars, (list, tuple)) else rebuilt[0]


def gradients(ys, xs, grad_ys=None, name=None, **kw):
    """Reverse-mode differentiation of sum(ys) with respect to the tensors `xs`."""
    ys = list(ys) if isinstance(ys, (list, tuple)) else [ys]
    xs = [convert_to_tensor(x) for x in xs]
    y = ys[0] if len(ys) == 1 else add_n(ys)
    # topological order of the sub-graph between xs and y
    order, seen = [], set()
    xset = {id(x) for x in xs}

    def visit(t):
        if id(t) in seen:
            return
        seen.add(id(t))
        if id(t) not in xset:
            for i in t.inputs:
                if isinstance(i, Tensor):
                    visit(i)
        order.append(t)
    visit(y)
    active = {}                     # does the tensor depend on any x?  (order is topological)
    nondiff = ("shape", "rank", "cmp", "not", "is_finite", "one_hot", "range", "zeros_like",
               "ones_like", "stop_gradient", "reduce_all")
    for t in order:
        cut = t.op in nondiff or t.op.startswith("assert_") or (t.op == "cast" and t.vjp is None)
        active[id(t)] = id(t) in xset or (not cut and any(
            isinstance(i, Tensor) and active.get(id(i), False) for i in t.inputs))
    cot = {id(y): ones_like(y)}
    for t in reversed(order):
        g = cot.get(id(t))
        if g is None or id(t) in xset or t.vjp is None:
            if g is not None and id(t) not in xset and t.vjp is None and active[id(t)] and \
                    t.op not in ("const", "variable", "placeholder", "loop_var", "zeros", "ones",
                                 "cmp", "shape", "range", "zeros_like", "ones_like",
                                 "stop_gradient"):
                raise NotImplementedError("tf.gradients through op %r (dtype %r, inputs %r)" % (
                    t.op, t._dtype, [(i.op, i._dtype) for i in t.inputs if isinstance(i, Tensor)]))
            continue
        if not active[id(t)]:
            continue
        for inp, gi in zip(t.inputs, t.vjp(g)):
            if gi is None or not isinstance(inp, Tensor):
                continue
            cot[id(inp)] = gi if id(inp) not in cot else add(cot[id(inp)], gi)
    return [cot.get(id(x)) for x in xs]


# ------------------------------------------------------------------------------------------------
_NOISE = {"normal": [], "uniform": []}


def set_noise(normal=(), uniform=()):
    """Arrays handed out, in evaluation order, by the tf.random_normal / tf.random_uniform ops of
    the next Session.run."""
    _NOISE["normal"] = list(normal)
    _NOISE["uniform"] = list(uniform)


def _noise_op(kind, shp, extra):
    def fn(c):
        want = _shape_arg(c, shp)
        if _is_peek(c):
            return extra(np.zeros(want, np.float32))
        if not _NOISE[kind]:
            raise RuntimeError("tf.random_%s evaluated but no injected noise is left" % kind)
        a = np.asarray(_NOISE[kind].pop(0), np.float32)
        assert tuple(a.shape) == want, (kind, a.shape, want)
        return extra(a)
    return Tensor(fn, op="random_" + kind, dtype=np.float32)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=np.float32, seed=None, name=None):   # noqa
    m, s = np.float32(mean), stddev
    if isinstance(s, Tensor):
        base = _noise_op("normal", shape, lambda a: a)
        return base * s + m if mean != 0.0 else base * s
    s = np.float32(s)
    return _noise_op("normal", shape, lambda a: (a * s + m) if (s != 1 or m != 0) else a)


def random_uniform(shape, minval=0, maxval=None, dtype=np.float32, seed=None, name=None):  # noqa
    return _noise_op("uniform", shape, lambda a: a)


class random(object):                                          # tf.random.*
    normal = staticmethod(random_normal)
    uniform = staticmethod(random_uniform)


def set_random_seed(seed):
    pass


# ------------------------------------------------------------------------------------------------
class Session(object):
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(dict(feed_dict or {}))

        def ev(f):
            if f is None:
                return None
            if isinstance(f, Tensor):
                v = ctx.eval(f)
                return None if v is None else np.array(v)
            if isinstance(f, dict):
                return {k: ev(v) for k, v in f.items()}
            if isinstance(f, (list, tuple)):
                out = [ev(x) for x in f]
                if hasattr(f, "_fields"):
                    return type(f)(*out)
                return out if isinstance(f, list) else tuple(out)
            for typ, fetch_fn in _RUN_CONVERSIONS:  # e.g. StochasticTensor (bn.py:311-316)
                if isinstance(f, typ):
                    tensors, contraction = fetch_fn(f)
                    return contraction([ev(t) for t in tensors])
            if hasattr(f, "__dict__"):              # plain result structs (HMCInfo)
                import copy
                o = copy.copy(f)
                o.__dict__ = {k: ev(v) for k, v in f.__dict__.items()}
                return o
            return f
        return ev(fetches)


class errors(object):
    class InvalidArgumentError(Exception):
        def __init__(self, node_def=None, op=None, message=""):
            Exception.__init__(self, message)
            self.message = message


class train(object):
    pass


# ---- what zhusuan/framework/bn.py and zhusuan/distributions/{base,utils,univariate}.py add -------
def rank(a, name=None):
    a = convert_to_tensor(a)
    return Tensor(lambda c: np.int32(np.ndim(c.eval(a))), inputs=(a,), op="rank", dtype=np.int32)


def _assert(check, what):
    def build(x, y=None, message=None, data=None, summarize=None, name=None):
        x = convert_to_tensor(x)
        y_t = None if y is None else (y if isinstance(y, Tensor) else None)

        def fn(c):
            yv = c.eval(y_t) if y_t is not None else y
            if not check(np.asarray(c.eval(x)), yv):
                raise errors.InvalidArgumentError(None, None, message or ("assert_%s failed" % what))
            return None
        return Tensor(fn, inputs=(x,) + ((y_t,) if y_t is not None else ()), op="assert_" + what)
    return build


assert_rank = _assert(lambda x, r: x.ndim == int(r), "rank")
assert_rank_at_least = _assert(lambda x, r: x.ndim >= int(r), "rank_at_least")
assert_greater_equal = _assert(lambda x, y: np.all(x >= y), "greater_equal")
assert_greater = _assert(lambda x, y: np.all(x > y), "greater")
assert_less_equal = _assert(lambda x, y: np.all(x <= y), "less_equal")
assert_positive = _assert(lambda x, y: np.all(x > 0), "positive")


def squeeze(a, axis=None, name=None, squeeze_dims=None):
    axis = squeeze_dims if axis is None else axis
    ax = None if axis is None else (tuple(axis) if isinstance(axis, (list, tuple)) else axis)
    a = convert_to_tensor(a)
    return Tensor(lambda c: np.squeeze(c.eval(a), axis=ax), inputs=(a,), op="squeeze",
                  vjp=lambda g: [reshape(g, shape(a))], dtype=a._dtype)


def concat(values, axis, name=None):
    vals = [convert_to_tensor(v) for v in values]
    ax_of = lambda c: int(c.eval(axis)) if isinstance(axis, Tensor) else int(axis)   # noqa: E731
    out = Tensor(lambda c: np.concatenate([np.atleast_1d(c.eval(v)) for v in vals], axis=ax_of(c)),
                 inputs=tuple(vals), op="concat", dtype=_dt(*vals))

    def vjp(g):
        def piece(k):
            def fn(c):
                sizes = [np.atleast_1d(c.eval(v)).shape[ax_of(c)] for v in vals]
                lo = int(np.sum(sizes[:k]))
                return np.take(np.asarray(c.eval(g)), np.arange(lo, lo + sizes[k]), axis=ax_of(c))
            return Tensor(fn, inputs=(g,) + tuple(vals), op="concat_grad", dtype=vals[k]._dtype)
        return [piece(k) for k in np.arange(len(vals))]
    out.vjp = vjp
    return out


def einsum(equation, *operands, **kw):
    """Two-operand einsum with explicit output (the only form the reference's examples use);
    the cotangents are einsums with the subscripts permuted."""
    a, b = [convert_to_tensor(o) for o in operands]
    lhs, res = equation.replace(" ", "").split("->")
    sa, sb = lhs.split(",")
    out = Tensor(lambda c: np.einsum(equation, c.eval(a), c.eval(b)), inputs=(a, b), op="einsum",
                 dtype=_dt(a, b))

    def vjp(g):
        def grad(own, other, other_t, own_t):
            def fn(c):
                r = np.einsum("%s,%s->%s" % (res, other, own), c.eval(g), c.eval(other_t))
                return r.astype(np.asarray(c.eval(own_t)).dtype, copy=False)
            return Tensor(fn, inputs=(g, other_t, own_t), op="einsum_grad", dtype=own_t._dtype)
        assert set(sa) <= set(res + sb) and set(sb) <= set(res + sa), "einsum vjp: summed-out index"
        return [grad(sa, sb, b, a), grad(sb, sa, a, b)]
    out.vjp = vjp
    return out


def reduce_prod(a, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    t = _reduce(np.prod, a, axis if axis is not None else reduction_indices,
                keepdims or (True if keep_dims else False), "prod", None)
    t.vjp = None                                   # not differentiated by the reference paths run
    return t


def reduce_all(a, axis=None, keepdims=False, name=None):
    a = convert_to_tensor(a)
    return Tensor(lambda c: np.all(c.eval(a), axis=axis, keepdims=keepdims), inputs=(a,),
                  op="reduce_all", dtype=np.bool_)


def reduce_max(a, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    axis = axis if axis is not None else reduction_indices
    kd = True if (keepdims or keep_dims) else False
    a = convert_to_tensor(a)
    t = _reduce(np.max, a, axis, kd, "max", None)
    t.vjp = _reduce_max_vjp(a, axis, kd, t)
    return t


def broadcast_static_shape(s1, s2):
    return TensorShape(np.broadcast_shapes(tuple(TensorShape(s1).as_list()),
                                           tuple(TensorShape(s2).as_list())))


def lgamma(a, name=None):
    from scipy.special import gammaln
    return _unary(lambda x: gammaln(x).astype(np.asarray(x).dtype), a, "lgamma")


class contrib(object):
    class distributions(object):
        pass


# ---- what ImportanceWeightedObjective.vimco (monte_carlo.py:166-227) adds -----------------------
def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=np.float32, name=None):
    def fn(c):
        i = np.asarray(c.eval(indices) if isinstance(indices, Tensor) else indices)
        d = int(c.eval(depth) if isinstance(depth, Tensor) else depth)
        return (np.arange(d) == i[..., None]).astype(dtype)
    ins = tuple(t for t in (indices, depth) if isinstance(t, Tensor))
    return Tensor(fn, inputs=ins, op="one_hot", dtype=dtype)


def transpose(a, perm=None, name=None):
    a = convert_to_tensor(a)

    def fn(c):
        p = None if perm is None else tuple(int(v) for v in np.asarray(
            c.eval(perm) if isinstance(perm, Tensor) else perm))
        return np.transpose(c.eval(a), p)

    def vjp(g):
        def gfn(c):
            p = None if perm is None else tuple(int(v) for v in np.asarray(
                c.eval(perm) if isinstance(perm, Tensor) else perm))
            return np.transpose(c.eval(g), None if p is None else tuple(np.argsort(p)))
        return [Tensor(gfn, inputs=(g,), op="transpose_grad", dtype=a._dtype)]
    ins = (a,) + ((perm,) if isinstance(perm, Tensor) else ())
    return Tensor(fn, inputs=ins, op="transpose", vjp=vjp, dtype=a._dtype)


def matrix_diag(a, name=None):
    a = convert_to_tensor(a)

    def fn(c):
        x = np.asarray(c.eval(a))
        return x[..., None] * np.eye(x.shape[-1], dtype=x.dtype)
    return Tensor(fn, inputs=(a,), op="matrix_diag", dtype=a._dtype)


# ---- what the VAE of examples/variational_autoencoders/iwae.py adds ------------------------------
def _reduce_max_vjp(a, axis, keepdims, out):
    def vjp(g):
        def gfn(c):
            x = np.asarray(c.eval(a))
            gv = np.asarray(c.eval(g))
            ax = _axes(c, axis, x.ndim)
            if ax is None:
                ax = tuple(np.arange(x.ndim))
            m = np.max(x, axis=ax, keepdims=True)
            if not keepdims:
                for d in sorted(ax):
                    gv = np.expand_dims(gv, d)
            ind = (x == m).astype(x.dtype)
            return (ind / ind.sum(axis=ax, keepdims=True) * gv).astype(x.dtype, copy=False)
        return [Tensor(gfn, inputs=(g, a), op="max_grad", dtype=a._dtype)]
    return vjp


def sigmoid(a, name=None):
    a = convert_to_tensor(a)
    f = lambda x: (1.0 / (1.0 + np.exp(-x))).astype(np.asarray(x).dtype)       # noqa: E731
    out = _unary(f, a, "sigmoid")
    out.vjp = lambda g: [g * out * (1.0 - out)]
    return out


def _dense_matmul(x, w):
    """x [..., in] . w [in, out] (tf.layers.dense contracts the last axis)."""
    x, w = convert_to_tensor(x), convert_to_tensor(w)
    out = Tensor(lambda c: np.matmul(c.eval(x), c.eval(w)), inputs=(x, w), op="dense_matmul",
                 dtype=_dt(x, w))

    def vjp(g):
        dx = Tensor(lambda c: np.matmul(c.eval(g), np.asarray(c.eval(w)).T), inputs=(g, w),
                    op="dense_dx", dtype=x._dtype)

        def dw(c):
            xv, gv = np.asarray(c.eval(x)), np.asarray(c.eval(g))
            return np.matmul(xv.reshape(-1, xv.shape[-1]).T, gv.reshape(-1, gv.shape[-1]))
        return [dx, Tensor(dw, inputs=(x, g), op="dense_dw", dtype=w._dtype)]
    out.vjp = vjp
    return out


_TEMPLATES = []          # stack of variable stores (tf.make_template / zs.reuse_variables)
_DEFAULT_STORE = {"vars": {}, "count": 0}
_TRAINABLE = []
_INIT = {"rng": None}


def set_init_rng(rng):
    """numpy Generator the Glorot-uniform initialiser of tf.layers.dense draws from."""
    _INIT["rng"] = rng


def trainable_variables(scope=None):
    return list(_TRAINABLE)


def reset_default_graph():
    del _TRAINABLE[:]
    _DEFAULT_STORE["vars"].clear()
    _DEFAULT_STORE["count"] = 0


class _Template(object):
    """tf.make_template: variables are created by the first call and reused by later ones."""

    def __init__(self, name, func):
        self.name, self.func = name, func
        self.store = {"vars": {}, "count": 0}

    def __call__(self, *args, **kwargs):
        self.store["count"] = 0
        _TEMPLATES.append(self.store)
        try:
            return self.func(*args, **kwargs)
        finally:
            _TEMPLATES.pop()


def make_template(name, func, create_scope_now_=False, unique_name_=None, custom_getter_=None,
                  **kwargs):
    return _Template(name, func)


class nn(object):
    @staticmethod
    def relu(a, name=None):
        a = convert_to_tensor(a)
        out = _unary(lambda x: np.maximum(x, 0).astype(np.asarray(x).dtype), a, "relu")
        out.vjp = lambda g: [g * cast(greater(a, 0.0), a._dtype or np.float32)]
        return out

    @staticmethod
    def sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        # max(x, 0) - x z + log(1 + exp(-|x|)), TF's numerically stable form
        z, x = convert_to_tensor(labels), convert_to_tensor(logits)

        def f(c):
            xv, zv = np.asarray(c.eval(x)), np.asarray(c.eval(z))
            return (np.maximum(xv, 0) - xv * zv + np.log1p(np.exp(-np.abs(xv)))).astype(xv.dtype)
        out = Tensor(f, inputs=(z, x), op="sigmoid_xent", dtype=x._dtype)
        out.vjp = lambda g: [_unbroadcast(negative(g * x), z), _unbroadcast(g * (sigmoid(x) - z), x)]
        return out

    sigmoid = staticmethod(sigmoid)

    @staticmethod
    def softmax(logits, axis=-1, name=None, dim=None):
        a = convert_to_tensor(logits)
        ax = dim if dim is not None else axis

        def f(x):
            e = np.exp(x - np.max(x, axis=ax, keepdims=True))
            return (e / e.sum(axis=ax, keepdims=True)).astype(np.asarray(x).dtype)
        out = _unary(f, a, "softmax")
        out.vjp = lambda g: [out * (g - reduce_sum(g * out, axis=ax, keepdims=True))]
        return out


class layers(object):
    @staticmethod
    def dense(inputs, units, activation=None, use_bias=True, name=None, **kw):
        store = _TEMPLATES[-1] if _TEMPLATES else _DEFAULT_STORE
        k = store["count"]
        store["count"] += 1
        key = name or ("dense" if k == 0 else "dense_%d" % k)
        inputs = convert_to_tensor(inputs)
        if key not in store["vars"]:
            fan_in = int(inputs.get_shape().as_list()[-1])
            limit = np.sqrt(6.0 / (fan_in + units))                 # glorot_uniform, TF's default
            w0 = _INIT["rng"].uniform(-limit, limit, (fan_in, units)).astype(np.float32)
            kern = Variable(w0, name=key + "/kernel")
            bias = Variable(np.zeros(units, np.float32), name=key + "/bias")
            store["vars"][key] = (kern, bias)
            _TRAINABLE.extend([kern, bias])
        kern, bias = store["vars"][key]
        y = _dense_matmul(inputs, kern)
        if use_bias:
            y = y + bias
        return activation(y) if activation is not None else y


# ---- `tensorflow.python.client.session` (zhusuan/framework/bn.py:10-11, variational/base.py:12) ---
_RUN_CONVERSIONS = []       # (type, fetch_function) registered by the reference's Tensor-likes
_TENSOR_CONVERSIONS = []    # (type, conversion function)


def register_tensor_conversion_function(base_type, conversion_func, priority=100):
    _TENSOR_CONVERSIONS.append((base_type, conversion_func))


def _register_session_run_conversion_functions(tensor_type, fetch_function, feed_function=None,
                                               feed_function_for_partial_run=None):
    _RUN_CONVERSIONS.append((tensor_type, fetch_function))


def constant_initializer(value=0, dtype=np.float32):
    return lambda shape=(): np.full(tuple(shape), value, dtype)


def zeros_initializer(dtype=np.float32):
    return lambda shape=(): np.zeros(tuple(shape), dtype)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    """Variables are keyed by name inside the innermost template (else globally): a second
    `get_variable` of the same name returns the same variable (the reference relies on that for
    REINFORCE's 'moving_mean', exclusive_kl.py:209-212)."""
    store = _TEMPLATES[-1] if _TEMPLATES else _DEFAULT_STORE
    key = "var:" + name
    if key not in store["vars"]:
        init = initializer if initializer is not None else zeros_initializer()
        v0 = init(tuple(shape or ())) if callable(init) else np.asarray(init)
        store["vars"][key] = Variable(np.asarray(v0, dtype or np.float32), name=name,
                                      trainable=trainable)
        if trainable:
            _TRAINABLE.append(store["vars"][key])
    return store["vars"][key]


def _assign_moving_average(variable, value, decay, zero_debias=True, name=None):
    """tensorflow/python/training/moving_averages.py (TF 1.13, the reference's pinned minimum:
    requirements-dev.txt:2), restated -- TensorFlow is a third-party dependency that is not in
    the reference checkout:

        d = 1 - decay
        zero_debias (the DEFAULT):  biased -= (biased - value) * d;  local_step += 1
                                    variable -= variable - biased / (1 - (1 - d) ** local_step)
        else:                       variable -= (variable - value) * d
    """
    d = np.float32(1.0) - np.float32(decay)
    value = convert_to_tensor(value)
    if not zero_debias:
        return assign(variable, variable - (variable - value) * d)
    if not hasattr(variable, "_zd"):
        variable._zd = (Variable(np.zeros_like(variable.value), name="biased", trainable=False),
                        Variable(np.zeros((), variable.value.dtype), name="local_step",
                                 trainable=False))
    biased, step = variable._zd
    upd_b = assign(biased, biased - (biased - value) * d)
    upd_s = assign(step, step + np.float32(1.0))
    return assign(variable, variable - (variable - upd_b / (np.float32(1.0)
                                                            - pow(np.float32(1.0) - d, upd_s))))


def _install_submodules():
    import sys
    import types
    me = sys.modules[__name__]
    me.__path__ = []                                    # importable as a package
    py = types.ModuleType(__name__ + ".python")
    client = types.ModuleType(__name__ + ".python.client")
    sess = types.ModuleType(__name__ + ".python.client.session")
    sess.register_session_run_conversion_functions = _register_session_run_conversion_functions
    training = types.ModuleType(__name__ + ".python.training")
    mavg = types.ModuleType(__name__ + ".python.training.moving_averages")
    mavg.assign_moving_average = _assign_moving_average
    training.moving_averages = mavg
    py.client, client.session, py.training = client, sess, training
    py.__path__, client.__path__, training.__path__ = [], [], []
    for m in (py, client, sess, training, mavg):
        sys.modules[m.__name__] = m
    me.python = py


_install_submodules()


def __getattr__(name):
    raise AttributeError("the NumPy TensorFlow shim (oracle/tf_shim) does not implement tf.%s"
                         % name)
