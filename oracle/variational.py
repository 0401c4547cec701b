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


def vimco_signal(log_w, axis, dtype=np.float32):
    """monte_carlo.py:194-223, restated the reference's way: an explicit
    [.., K, K] tile whose row k is log_w with entry k replaced by the mean of
    the others, reduced with log_mean_exp.  Returns LME(log_w, keepdims) - cv."""
    l = np.asarray(log_w, dtype)
    K = l.shape[axis]
    if K < 2:
        raise ValueError(
            "VIMCO is a multi-sample gradient estimator, size along "
            "`axis` in the objective should be larger than 1.")
    mean_except = (l.sum(axis=axis, keepdims=True, dtype=dtype) - l) / dtype(K - 1)
    x = np.moveaxis(l, axis, -1)                 # transpose(perm): axis <-> last
    sub = np.moveaxis(mean_except, axis, -1)
    x_ex = np.repeat(x[..., None], K, axis=-1)   # tile: x_ex[..., j, k] = x[..., j]
    idx = np.arange(K)
    x_ex[..., idx, idx] = sub                    # - diag(x) + diag(sub_x)
    cv = log_mean_exp(np.swapaxes(x_ex, -1, -2), axis=-1, dtype=dtype)
    cv = np.moveaxis(cv, -1, axis)
    return (log_mean_exp(l, axis, keepdims=True, dtype=dtype) - cv).astype(dtype)


def vimco_cost(log_joint, log_q, axis, dtype=np.float32):
    """monte_carlo.py:221-227 (single latent): fake_term = sum(log q * signal)."""
    log_w = np.asarray(log_joint, dtype) - np.asarray(log_q, dtype)
    sig = vimco_signal(log_w, axis, dtype)
    fake = (np.asarray(log_q, dtype) * sig).sum(axis=axis, dtype=dtype)
    return (-fake - log_mean_exp(log_w, axis, dtype=dtype)).astype(dtype)


def vimco_grad_logq(log_w, axis, dtype=np.float32):
    """d sum(vimco_cost) / d log q (samples held fixed): -signal + softmax."""
    return (-vimco_signal(log_w, axis, dtype) + iw_grad_logw(log_w, axis, dtype)).astype(dtype)


def normalized_weights(log_w, axis, dtype=np.float32):
    """inclusive_kl.py:139-143."""
    l = np.asarray(log_w, dtype)
    w_u = np.exp(l - l.max(axis=axis, keepdims=True))
    return (w_u / w_u.sum(axis=axis, keepdims=True, dtype=dtype)).astype(dtype)


def importance_cost(log_joint, log_q, axis, dtype=np.float32):
    """inclusive_kl.py:137-151: sum_axis(w~ * (-log q))."""
    log_q = np.asarray(log_q, dtype)
    w = normalized_weights(np.asarray(log_joint, dtype) - log_q, axis, dtype)
    return (w * (-log_q)).sum(axis=axis, dtype=dtype).astype(dtype)


def reinforce_cost(log_joint, log_q, axis, dtype=np.float32):
    """exclusive_kl.py:213-225 with variance_reduction=False (single latent):
    cost = mean_axis(-log p + stop_gradient(log p - log q) * (-log q))."""
    lj, lq = np.asarray(log_joint, dtype), np.asarray(log_q, dtype)
    cost = -lj + (lj - lq) * (-lq)
    return cost.mean(axis=axis, dtype=dtype) if axis is not None else cost


def reinforce_grad_logq(log_joint, log_q, axis, dtype=np.float32):
    """d sum(reinforce_cost) / d log q with the signal held constant."""
    lj, lq = np.asarray(log_joint, dtype), np.asarray(log_q, dtype)
    k = lq.shape[axis] if axis is not None else 1
    return (-(lj - lq) / dtype(k)).astype(dtype)


def zero_debiased_moving_average(state, value, decay):
    """``moving_averages.assign_moving_average(var, value, decay)`` of TensorFlow 1.x with its
    default ``zero_debias=True`` (tensorflow/python/training/moving_averages.py; TF is a third-party
    dependency absent from the reference checkout -- pinned version: requirements-dev.txt:2), as
    exclusive_kl.py:215-216 calls it for REINFORCE's baseline:
        biased -= (biased - value) * (1 - decay);  step += 1;  var = biased / (1 - decay ** step)
    state = (biased, step) or None.  Returns (var, state)."""
    biased, step = state if state is not None else (0.0, 0)
    biased = biased - (biased - value) * (1.0 - decay)
    step += 1
    return biased / (1.0 - decay ** step), (biased, step)
