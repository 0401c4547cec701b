"""TensorArithmeticMixin -- the Tensor-like behaviour StochasticTensor and the
variational objectives rely on (zhusuan/utils.py:23-150) -- with the data of the
reference's own tests (tests/test_utils.py:47-255): unary / binary / relational /
logical operators on either side, __getitem__, and the disallowed iter / bool."""
import operator

import numpy as np
import pytest
import torch

from zhusuan_b200.utils import TensorArithmeticMixin


class _SimpleTensor(TensorArithmeticMixin):
    def __init__(self, value):
        self.value = torch.as_tensor(value)

    @property
    def tensor(self):
        return self.value

    dtype = property(lambda self: self.value.dtype)
    shape = property(lambda self: self.value.shape)


def _check(func, *np_args):
    """func on plain tensors vs func with each / all operands wrapped."""
    ts = [torch.as_tensor(a) for a in np_args]
    ans = func(*ts)
    n = len(ts)
    variants = [[_SimpleTensor(t) if (mask >> i) & 1 else t for i, t in enumerate(ts)]
                for mask in range(1, 1 << n)]
    for args in variants:
        res = func(*args)
        res = res.tensor if isinstance(res, TensorArithmeticMixin) else res
        assert res.dtype == ans.dtype
        np.testing.assert_array_equal(res.numpy(), ans.numpy())


def test_unary_ops():
    int_data = np.asarray([1, -2, 3], dtype=np.int32)
    float_data = np.asarray([1.1, -2.2, 3.3], dtype=np.float32)
    bool_data = np.asarray([True, False, True])
    for d in (int_data, float_data):
        _check(abs, d)
        _check(operator.neg, d)
    _check(operator.invert, bool_data)


def test_binary_ops():
    arith = [operator.add, operator.sub, operator.mul, operator.truediv, operator.floordiv,
             operator.mod]
    xi, yi = np.asarray([-4, 5, 6], np.int32), np.asarray([1, -2, 3], np.int32)
    xf, yf = np.asarray([-4.4, 5.5, 6.6], np.float32), np.asarray([1.1, -2.2, 3.3], np.float32)
    for op in arith:
        _check(op, xi, yi)
        _check(op, xf, yf)
    _check(operator.pow, xi, np.asarray([1, 2, 3], np.int32))
    _check(operator.pow, np.abs(xf), yf)
    xb = np.asarray([True, False, True, False])
    yb = np.asarray([True, True, False, False])
    for op in (operator.and_, operator.or_, operator.xor):
        _check(op, xb, yb)
    a = np.asarray([1, -2, 3, -4, 5, 6, -4, 5, 6], np.int32)
    b = np.asarray([1, -2, 3, 1, -2, 3, -4, 5, 6], np.int32)
    for op in (operator.lt, operator.le, operator.gt, operator.ge):
        _check(op, a, b)
        _check(op, a.astype(np.float32) * 1.1, b.astype(np.float32) * 1.1)


def test_getitem():
    data = np.asarray([1, 2, 3, 4, 5, 6, 7, 8], dtype=np.int32)
    x = _SimpleTensor(data)
    for s in [0, -1, slice(0, None), slice(None, 1), slice(None, None, 2), slice(-1, None),
              slice(None, -1)]:
        res = x[s]
        res = res.tensor if isinstance(res, TensorArithmeticMixin) else res
        np.testing.assert_array_equal(res.numpy(), data[s])
    np.testing.assert_array_equal(x[torch.tensor(3)].numpy(), data[3])


def test_disallowed_operators():
    with pytest.raises(TypeError, match="_SimpleTensor object is not iterable"):
        iter(_SimpleTensor(1))
    with pytest.raises(TypeError, match="Using a `_SimpleTensor` object as a Python `bool` "
                                        "is not allowed"):
        not _SimpleTensor(1)
    with pytest.raises(TypeError, match="as a Python `bool` is not allowed"):
        if _SimpleTensor(1):
            pass


def test_torch_functions_accept_the_wrapper():
    """the analogue of tf.register_tensor_conversion_function: torch.* calls unwrap."""
    s = _SimpleTensor(np.asarray([1.0, 4.0], np.float32))
    np.testing.assert_allclose(torch.sqrt(s).numpy(), [1.0, 2.0])
    assert float(torch.sum(s * 2)) == 10.0
    assert torch.stack([s, s]).shape == (2, 2)
