"""Host-side helpers mirroring zhusuan/utils.py (the tensor-like mixin,
``log_mean_exp`` / ``log_sum_exp``, ``merge_dicts``)."""
import torch

from . import ops

__all__ = ["TensorArithmeticMixin", "log_mean_exp", "log_sum_exp",
           "merge_dicts", "convert_to_tensor"]


def convert_to_tensor(x, dtype=None, device=None):
    """``tf.convert_to_tensor`` for the torch world; tensor-likes
    (StochasticTensor, VariationalObjective) expose ``.tensor``."""
    if isinstance(x, TensorArithmeticMixin):
        x = x.tensor
    if isinstance(x, torch.Tensor):
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        if device is not None and x.device != torch.device(device):
            x = x.to(device)
        return x
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if dtype is None and isinstance(x, float):
        dtype = torch.float32
    if isinstance(x, (bool, int, float)):
        # a device-side fill, not a host-to-device copy: legal inside CUDA-graph capture
        dt = dtype or (torch.bool if isinstance(x, bool) else
                       torch.int32 if isinstance(x, int) else torch.float32)
        return torch.full((), x, dtype=dt, device=device)
    t = torch.as_tensor(x, device=device)
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype == torch.float64:
        t = t.to(torch.float32)      # TF's default float is float32
    elif t.dtype == torch.int64:
        t = t.to(torch.int32)
    return t


class TensorArithmeticMixin(object):
    """zhusuan/utils.py:18-150: objects with a ``.tensor`` behave like that
    tensor in arithmetic and in ``torch.*`` calls."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def unwrap(a):
            if isinstance(a, TensorArithmeticMixin):
                return a.tensor
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(b) for b in a)
            return a
        return func(*[unwrap(a) for a in args],
                    **{k: unwrap(v) for k, v in (kwargs or {}).items()})

    def _t(self):
        return self.tensor

    def __abs__(self): return abs(self._t())
    def __neg__(self): return -self._t()
    def __add__(self, o): return self._t() + _u(o)
    def __radd__(self, o): return _u(o) + self._t()
    def __sub__(self, o): return self._t() - _u(o)
    def __rsub__(self, o): return _u(o) - self._t()
    def __mul__(self, o): return self._t() * _u(o)
    def __rmul__(self, o): return _u(o) * self._t()
    def __truediv__(self, o): return self._t() / _u(o)
    def __rtruediv__(self, o): return _u(o) / self._t()
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __floordiv__(self, o): return self._t() // _u(o)
    def __rfloordiv__(self, o): return _u(o) // self._t()
    def __mod__(self, o): return self._t() % _u(o)
    def __rmod__(self, o): return _u(o) % self._t()
    def __pow__(self, o): return self._t() ** _u(o)
    def __rpow__(self, o): return _u(o) ** self._t()
    def __matmul__(self, o): return self._t() @ _u(o)
    def __rmatmul__(self, o): return _u(o) @ self._t()
    # logical operators (zhusuan/utils.py:95-117)
    def __invert__(self): return ~self._t()
    def __and__(self, o): return self._t() & _u(o)
    def __rand__(self, o): return _u(o) & self._t()
    def __or__(self, o): return self._t() | _u(o)
    def __ror__(self, o): return _u(o) | self._t()
    def __xor__(self, o): return self._t() ^ _u(o)
    def __rxor__(self, o): return _u(o) ^ self._t()
    def __lt__(self, o): return self._t() < _u(o)
    def __le__(self, o): return self._t() <= _u(o)
    def __gt__(self, o): return self._t() > _u(o)
    def __ge__(self, o): return self._t() >= _u(o)
    def __getitem__(self, item): return self._t()[item]
    def __hash__(self): return id(self)
    def __eq__(self, o): return id(self) == id(o)

    def __iter__(self):
        raise TypeError("{} object is not iterable.".format(
            self.__class__.__name__))

    def __bool__(self):
        raise TypeError(
            "Using a `{}` object as a Python `bool` is not allowed. "
            "Use `if t is not None:` instead of `if t:` to test if a "
            "tensor is defined.".format(self.__class__.__name__))
    __nonzero__ = __bool__


def _u(o):
    return o.tensor if isinstance(o, TensorArithmeticMixin) else o


def log_sum_exp(x, axis=None, keepdims=False):
    """zhusuan/utils.py:153-174 (kernel: zsb_reduce_fwd_f32 op 2)."""
    return ops.reduce_axes(convert_to_tensor(x, torch.float32), ops.OP_LSE,
                           axis, keepdims)


def log_mean_exp(x, axis=None, keepdims=False):
    """zhusuan/utils.py:177-196 (kernel: zsb_reduce_fwd_f32 op 0)."""
    return ops.reduce_axes(convert_to_tensor(x, torch.float32), ops.OP_LME,
                           axis, keepdims)


def merge_dicts(*dict_args):
    """zhusuan/utils.py:220-228."""
    result = {}
    for d in dict_args:
        result.update(d)
    return result
