"""dtype / shape assertion helpers (zhusuan/distributions/utils.py:80-210)."""
import torch

FLOATS = [torch.float16, torch.float32, torch.float64]
INTS = [torch.int16, torch.int32, torch.int64]


def assert_same_dtype_in(tensors_with_name, dtypes=None):
    expected = None
    for tensor, name in tensors_with_name:
        if dtypes and tensor.dtype not in dtypes:
            if len(dtypes) == 1:
                raise TypeError('{}({}) must have dtype {}.'.format(
                    name, tensor.dtype, dtypes[0]))
            raise TypeError('{}({}) must have a dtype in {}.'.format(
                name, tensor.dtype, dtypes))
        if expected is None:
            expected = tensor.dtype
        elif expected != tensor.dtype:
            t0, n0 = tensors_with_name[0]
            raise TypeError('{}({}) must have the same dtype as {}({}).'
                            .format(name, tensor.dtype, n0, t0.dtype))
    return expected


def assert_same_float_dtype(tensors_with_name):
    return assert_same_dtype_in(tensors_with_name, FLOATS)


def assert_dtype_in_dtypes(dtype, dtypes):
    if dtype not in dtypes:
        raise TypeError("`dtype`({}) not in {}".format(dtype, dtypes))


def assert_dtype_is_int_or_float(dtype):
    assert_dtype_in_dtypes(dtype, INTS + FLOATS)


def assert_rank_at_least(tensor, k, name):
    if tensor.dim() < k:
        raise ValueError('{} should have rank >= {}.'.format(name, k))
    return tensor


def broadcast_check(a, b, msg):
    try:
        return torch.broadcast_shapes(tuple(a), tuple(b))
    except RuntimeError:
        raise ValueError(msg)
