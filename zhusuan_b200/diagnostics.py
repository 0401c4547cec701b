"""Chain diagnostics on the device (zhusuan/diagnostics.py:17-64).

The reference computes the Stan effective-sample-size estimator with NumPy on
samples fetched to the host; here the [M, D] sample matrix stays in HBM and one
kernel (zsb_effective_sample_size_f32) produces the per-dimension estimates.
"""
import numpy as np
import torch

from ._lib import lib, ptr, stream

__all__ = ['effective_sample_size', 'effective_sample_size_1d',
           'effective_sample_size_per_dim']


def _as_device_matrix(samples):
    if isinstance(samples, np.ndarray):
        samples = torch.from_numpy(np.ascontiguousarray(samples))
    if not isinstance(samples, torch.Tensor):
        raise TypeError("samples must be a numpy array or a torch tensor")
    if not samples.is_cuda:
        samples = samples.cuda()          # no CPU implementation: the kernel is the product
    return samples.to(torch.float32).contiguous()


def effective_sample_size_per_dim(samples, burn_in=100):
    """[M, D] -> device tensor [D] of per-dimension effective sample sizes."""
    s = _as_device_matrix(samples)
    if s.dim() != 2:
        raise ValueError("samples should be a 2-D array of shape (M, D)")
    s = s[burn_in:].contiguous()
    M, D = int(s.shape[0]), int(s.shape[1])
    out = torch.empty(D, dtype=torch.float32, device=s.device)
    lib.call("zsb_effective_sample_size_f32", ptr(s), M, D, ptr(out), stream())
    return out


def effective_sample_size_1d(samples):
    """diagnostics.py:17-41: a 1-D chain of scalar samples -> float."""
    s = _as_device_matrix(samples).reshape(-1, 1)
    return float(effective_sample_size_per_dim(s, burn_in=0)[0])


def effective_sample_size(samples, burn_in=100):
    """diagnostics.py:44-64: the minimum positive per-dimension estimate
    (``inf`` if none is positive)."""
    ess = effective_sample_size_per_dim(samples, burn_in).double().cpu().numpy()
    assert (ess >= 0).all()
    pos = ess[ess > 0]
    return float(pos.min()) if pos.size else float("inf")
