"""NumPy restatement of the multi-sample objectives (TEST ORACLE ONLY).

  log_mean_exp            zhusuan/utils.py:177-196
  ELBO  (.tensor/.sgvb)   zhusuan/variational/exclusive_kl.py:131-159
  IWAE  (.tensor/.sgvb)   zhusuan/variational/monte_carlo.py:137-164
  VariationalObjective    zhusuan/variational/base.py:169-183
      log_w = log_joint + entropy,  entropy = -sum_z log q(z)

The "reparameterised gradient" is TF autodiff through these reductions; the
backward weights are restated analytically here:
  d mean_axis(x) / dx      = 1/K
  d log_mean_exp(x) / dx   = softmax_axis(x)
"""
import numpy as np


def log_mean_exp(x, axis=None, keepdims=False, dtype=np.float32):
    """zhusuan/utils.py:190-196."""
    x = np.asarray(x, dtype)
    x_max = x.max(axis=axis, keepdims=True)
    ret = np.log(np.mean(np.exp(x - x_max), axis=axis, keepdims=True,
                         dtype=dtype)) + x_max
    if not keepdims:
        ret = ret.mean(axis=axis, dtype=dtype)   # utils.py:195: reduce_mean
    return ret.astype(dtype)


def log_sum_exp(x, axis=None, keepdims=False, dtype=np.float32):
    """zhusuan/utils.py:153-174."""
    x = np.asarray(x, dtype)
    x_max = x.max(axis=axis, keepdims=True)
    ret = np.log(np.sum(np.exp(x - x_max), axis=axis, keepdims=True,
                        dtype=dtype)) + x_max
    if not keepdims:
        ret = ret.sum(axis=axis, dtype=dtype)
    return ret.astype(dtype)


def elbo(log_joint, log_qs, axis=None, dtype=np.float32):
    """exclusive_kl.py:131-137.  log_qs: list of log q(z) terms (may be [])."""
    lb = np.asarray(log_joint, dtype)
    for lq in log_qs:
        lb = lb - np.asarray(lq, dtype)
    if axis is not None:
        lb = lb.mean(axis=axis, dtype=dtype)
    return lb.astype(dtype)


def elbo_grad_logw(shape, axis, dtype=np.float32):
    """d sum(elbo) / d log_w: uniform 1/K along ``axis``."""
    k = shape[axis] if axis is not None else 1
    return np.full(shape, dtype(1.0) / dtype(k), dtype)


def iw_objective(log_joint, log_qs, axis, dtype=np.float32):
    """monte_carlo.py:137-141."""
    if axis is None:
        raise ValueError(
            "ImportanceWeightedObjective is a multi-sample objective, "
            "the `axis` argument must be specified.")
    log_w = np.asarray(log_joint, dtype)
    for lq in log_qs:
        log_w = log_w - np.asarray(lq, dtype)
    return log_mean_exp(log_w, axis, dtype=dtype)


def iw_grad_logw(log_w, axis, dtype=np.float32):
    """d sum(log_mean_exp(log_w, axis)) / d log_w = softmax over ``axis``."""
    log_w = np.asarray(log_w, dtype)
    m = log_w.max(axis=axis, keepdims=True)
    e = np.exp(log_w - m)
    return (e / e.sum(axis=axis, keepdims=True, dtype=dtype)).astype(dtype)
