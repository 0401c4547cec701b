"""NumPy restatement of the hot-path ``log_prob`` formulas (TEST ORACLE ONLY).

Each function follows the reference's ``_log_prob`` followed by
``Distribution.log_prob``'s reduce_sum over the last ``group_ndims`` axes
(zhusuan/distributions/base.py:290-304).  Broadcasting is NumPy broadcasting,
which is what ``maybe_explicit_broadcast`` materialises
(zhusuan/distributions/utils.py:52-78).

``dtype`` selects the arithmetic type (float32 mirrors TF's default, float64
is the high-precision twin used to bound fp32 error in the parity tests).
"""
import numpy as np
from scipy import special as _sp

LOG_2PI_HALF = 0.5 * np.log(2.0 * np.pi)


def _group_sum(x, group_ndims):
    # distributions/base.py:303-304: reduce_sum(log_p, range(-group_ndims, 0))
    if group_ndims == 0:
        return x
    return x.sum(axis=tuple(range(-group_ndims, 0)))


def normal_log_prob(given, mean, logstd, group_ndims=0, dtype=np.float32):
    """Normal._log_prob, univariate.py:174-181.

    c - logstd - 0.5 * exp(-2 logstd) * (given - mean)^2, c = -0.5 log(2 pi).
    """
    given, mean, logstd = (np.asarray(a, dtype) for a in (given, mean, logstd))
    c = dtype(-LOG_2PI_HALF)
    precision = np.exp(dtype(-2.0) * logstd)
    lp = c - logstd - dtype(0.5) * precision * np.square(given - mean)
    return _group_sum(lp, group_ndims)


def normal_log_prob_grads(given, mean, logstd, dtype=np.float64):
    """Elementwise d log_prob / d(given, mean, logstd) (for the K1 bwd test)."""
    given, mean, logstd = (np.asarray(a, dtype) for a in (given, mean, logstd))
    prec = np.exp(-2.0 * logstd)
    diff = given - mean
    dgiven = -prec * diff
    dmean = prec * diff
    dlogstd = -1.0 + prec * diff * diff
    return dgiven, dmean, dlogstd


def bernoulli_log_prob(given, logits, group_ndims=0, dtype=np.float32):
    """Bernoulli._log_prob, univariate.py:398-403.

    -sigmoid_cross_entropy_with_logits(labels=x, logits=l)
      = -(max(l, 0) - l * x + log(1 + exp(-|l|)))   (TF's stable form).
    """
    x = np.asarray(given).astype(dtype)
    l = np.asarray(logits, dtype)
    lp = -(np.maximum(l, dtype(0)) - l * x + np.log1p(np.exp(-np.abs(l))))
    return _group_sum(lp.astype(dtype), group_ndims)


def categorical_log_prob(given, logits, group_ndims=0, dtype=np.float32):
    """Categorical._log_prob, univariate.py:496-548.

    -sparse_softmax_cross_entropy_with_logits = log_softmax(logits)[given].
    """
    logits = np.asarray(logits, dtype)
    given = np.asarray(given).astype(np.int64)
    bshape = np.broadcast_shapes(given.shape, logits.shape[:-1])
    given = np.broadcast_to(given, bshape)
    logits = np.broadcast_to(logits, bshape + logits.shape[-1:])
    m = logits.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(logits - m).sum(axis=-1, keepdims=True)) + m
    logsm = logits - lse
    lp = np.take_along_axis(logsm, given[..., None], axis=-1)[..., 0]
    return _group_sum(lp.astype(dtype), group_ndims)


def dirichlet_log_prob(given, alpha, group_ndims=0, dtype=np.float32):
    """Dirichlet._log_prob, multivariate.py:665-677.

    -lbeta(alpha) + sum((alpha - 1) * log(given), -1),
    lbeta = sum(lgamma(alpha_i)) - lgamma(sum(alpha_i)).
    """
    given = np.asarray(given, dtype)
    alpha = np.asarray(alpha, dtype)
    given, alpha = np.broadcast_arrays(given, alpha)
    lbeta = _sp.gammaln(alpha).sum(-1) - _sp.gammaln(alpha.sum(-1))
    lp = -lbeta + ((alpha - dtype(1)) * np.log(given)).sum(-1)
    return _group_sum(lp.astype(dtype), group_ndims)


def unnormalized_multinomial_log_prob(given, logits, normalize_logits=True,
                                      group_ndims=0, dtype=np.float32):
    """UnnormalizedMultinomial._log_prob, multivariate.py:435-443."""
    given = np.asarray(given).astype(dtype)
    logits = np.asarray(logits, dtype)
    given, logits = np.broadcast_arrays(given, logits)
    if normalize_logits:
        m = logits.max(axis=-1, keepdims=True)
        logits = logits - (np.log(np.exp(logits - m).sum(-1, keepdims=True))
                           + m)
    lp = (given * logits).sum(-1)
    return _group_sum(lp.astype(dtype), group_ndims)


def mvn_cholesky_log_prob(given, mean, cov_tril, group_ndims=0,
                          dtype=np.float32):
    """MultivariateNormalCholesky._log_prob, multivariate.py:169-189.

    log_z = -n/2 log(2 pi) - sum(log diag L);  -0.5 |L^{-1}(given - mean)|^2.
    """
    from scipy.linalg import solve_triangular
    given = np.asarray(given, dtype)
    mean = np.asarray(mean, dtype)
    L = np.asarray(cov_tril, dtype)
    n = mean.shape[-1]
    log_det = 2 * np.log(np.diagonal(L, axis1=-2, axis2=-1)).sum(-1)
    log_z = -n / 2 * np.log(2 * np.pi) - log_det / 2
    y = given - mean
    bshape = np.broadcast_shapes(y.shape[:-1], L.shape[:-2])
    y = np.broadcast_to(y, bshape + (n,)).reshape(-1, n)
    Lb = np.broadcast_to(L, bshape + (n, n)).reshape(-1, n, n)
    x = np.stack([solve_triangular(Lb[i], y[i], lower=True)
                  for i in range(y.shape[0])]) if y.shape[0] else \
        np.zeros((0, n), dtype)
    stoc = -0.5 * np.square(x).sum(-1).reshape(bshape)
    lp = (np.broadcast_to(log_z, bshape) + stoc).astype(dtype)
    return _group_sum(lp, group_ndims)


def normal_sample(eps, mean, std, dtype=np.float32):
    """Normal._sample with injected eps: eps * std + mean (univariate.py:167)."""
    return (np.asarray(eps, dtype) * np.asarray(std, dtype)
            + np.asarray(mean, dtype))


def bernoulli_sample(u, logits, dtype=np.int32):
    """Bernoulli._sample with injected uniforms: u < sigmoid(logits)
    (univariate.py:386-392)."""
    p = 1.0 / (1.0 + np.exp(-np.asarray(logits, np.float32)))
    return (np.asarray(u, np.float32) < p.astype(np.float32)).astype(dtype)


# ---- the other elementwise univariate families (zhusuan/distributions/univariate.py) ---------
def _softplus(t):
    return np.maximum(t, 0) + np.log1p(np.exp(-np.abs(t)))


def fold_normal_log_prob(given, mean, logstd, group_ndims=0, dtype=np.float32):
    """univariate.py:319-329."""
    x, m, ls = (np.asarray(v, dtype) for v in (given, mean, logstd))
    c = dtype(-0.5 * (np.log(2.0) + np.log(np.pi)))
    prec = np.exp(-2 * ls)
    with np.errstate(divide="ignore"):
        mask = np.log((x >= 0).astype(dtype))
    lp = (c - (ls + 0.5 * prec * np.square(x - m)) + _softplus(-2 * m * x * prec)) + mask
    return _group_sum(lp.astype(dtype), group_ndims)


def uniform_log_prob(given, minval, maxval, group_ndims=0, dtype=np.float32):
    """univariate.py:646-660: log(1/(max-min) * [min <= x < max])."""
    x, lo, hi = (np.asarray(v, dtype) for v in (given, minval, maxval))
    mask = np.logical_and(lo <= x, x < hi).astype(dtype)
    with np.errstate(divide="ignore"):
        lp = np.log(1 / (hi - lo) * mask)
    return _group_sum(lp.astype(dtype), group_ndims)


def gamma_log_prob(given, alpha, beta, group_ndims=0, dtype=np.float32):
    """univariate.py:737-747."""
    x, a, b = (np.asarray(v, dtype) for v in (given, alpha, beta))
    lp = a * np.log(b) - _sp.gammaln(a) + (a - 1) * np.log(x) - b * x
    return _group_sum(lp.astype(dtype), group_ndims)


def beta_log_prob(given, alpha, beta, group_ndims=0, dtype=np.float32):
    """univariate.py:833-851."""
    x, a, b = (np.asarray(v, dtype) for v in (given, alpha, beta))
    lp = (a - 1) * np.log(x) + (b - 1) * np.log(1 - x) - (
        _sp.gammaln(a) + _sp.gammaln(b) - _sp.gammaln(a + b))
    return _group_sum(lp.astype(dtype), group_ndims)


def poisson_log_prob(given, rate, group_ndims=0, dtype=np.float32):
    """univariate.py:922-933."""
    x, r = np.asarray(given, dtype), np.asarray(rate, dtype)
    lp = x * np.log(r) - r - _sp.gammaln(x + 1)
    return _group_sum(lp.astype(dtype), group_ndims)


def binomial_log_prob(given, logits, n_experiments, group_ndims=0, dtype=np.float32):
    """univariate.py:1047-1064."""
    x, l = np.asarray(given, dtype), np.asarray(logits, dtype)
    n = dtype(n_experiments)
    lp = _sp.gammaln(n + 1) - _sp.gammaln(n - x + 1) - _sp.gammaln(x + 1) + x * l + \
        n * (-_softplus(l))
    return _group_sum(lp.astype(dtype), group_ndims)


def inverse_gamma_log_prob(given, alpha, beta, group_ndims=0, dtype=np.float32):
    """univariate.py:1146-1158."""
    x, a, b = (np.asarray(v, dtype) for v in (given, alpha, beta))
    lp = a * np.log(b) - _sp.gammaln(a) - (a + 1) * np.log(x) - b / x
    return _group_sum(lp.astype(dtype), group_ndims)


def laplace_log_prob(given, loc, scale, group_ndims=0, dtype=np.float32):
    """univariate.py:1267-1273."""
    x, m, s = (np.asarray(v, dtype) for v in (given, loc, scale))
    lp = -np.log(dtype(2.)) - np.log(s) - np.abs(x - m) / s
    return _group_sum(lp.astype(dtype), group_ndims)


def bin_concrete_log_prob(given, temperature, logits, group_ndims=0, dtype=np.float32):
    """univariate.py:1381-1400."""
    x, t, l = (np.asarray(v, dtype) for v in (given, temperature, logits))
    lx, l1x = np.log(x), np.log(1 - x)
    temp = t * (lx - l1x) - l
    lp = np.log(t) - lx - l1x + temp - 2 * _softplus(temp)
    return _group_sum(lp.astype(dtype), group_ndims)


def multinomial_log_prob(given, logits, n_experiments=None, normalize_logits=True,
                         group_ndims=0, dtype=np.float32):
    """multivariate.py:313-331: log n! - sum log k_i! + sum k_i * logits_i (normalised)."""
    g, l = np.broadcast_arrays(np.asarray(given, dtype), np.asarray(logits, dtype))
    if normalize_logits:
        m = l.max(-1, keepdims=True)
        l = l - (np.log(np.exp(l - m).sum(-1, keepdims=True)) + m)
    n = g.sum(-1) if n_experiments is None else dtype(n_experiments)
    lp = _sp.gammaln(n + 1) - _sp.gammaln(g + 1).sum(-1) + (g * l).sum(-1)
    return _group_sum(lp.astype(dtype), group_ndims)


def onehot_categorical_log_prob(given, logits, group_ndims=0, dtype=np.float32):
    """multivariate.py:517-536: -softmax_cross_entropy(labels=given, logits)."""
    g, l = np.broadcast_arrays(np.asarray(given, dtype), np.asarray(logits, dtype))
    m = l.max(-1, keepdims=True)
    l = l - (np.log(np.exp(l - m).sum(-1, keepdims=True)) + m)
    return _group_sum((g * l).sum(-1).astype(dtype), group_ndims)


def exp_concrete_log_prob(given, temperature, logits, group_ndims=0, dtype=np.float32):
    """multivariate.py:800-812."""
    x, l = np.asarray(given, dtype), np.asarray(logits, dtype)
    t = dtype(temperature)
    n = l.shape[-1]
    temp = l - t * x
    m = temp.max(-1, keepdims=True)
    lse = (np.log(np.exp(temp - m).sum(-1, keepdims=True)) + m)[..., 0]
    lp = _sp.gammaln(n) + (n - 1) * np.log(t) + temp.sum(-1) - n * lse
    return _group_sum(lp.astype(dtype), group_ndims)


def concrete_log_prob(given, temperature, logits, group_ndims=0, dtype=np.float32):
    """multivariate.py:938-955."""
    x, l = np.asarray(given, dtype), np.asarray(logits, dtype)
    t = dtype(temperature)
    n = l.shape[-1]
    lx = np.log(x)
    temp = l - t * lx
    m = temp.max(-1, keepdims=True)
    lse = (np.log(np.exp(temp - m).sum(-1, keepdims=True)) + m)[..., 0]
    lp = _sp.gammaln(n) + (n - 1) * np.log(t) + (temp - lx).sum(-1) - n * lse
    return _group_sum(lp.astype(dtype), group_ndims)


def matrix_normal_cholesky_log_prob(given, mean, u_tril, v_tril, dtype=np.float64):
    """multivariate.py:1123-1157 for a single (unbatched) distribution; given [..., r, c]."""
    from scipy.linalg import solve_triangular
    y = np.asarray(given, dtype) - np.asarray(mean, dtype)
    lu, lv = np.asarray(u_tril, dtype), np.asarray(v_tril, dtype)
    r, c = y.shape[-2:]
    log_z = -(r * c) / 2. * np.log(2 * np.pi) - r / 2. * 2 * np.log(np.diag(lv)).sum() \
        - c / 2. * 2 * np.log(np.diag(lu)).sum()
    out = np.empty(y.shape[:-2], dtype)
    for idx in np.ndindex(*y.shape[:-2]):
        a = solve_triangular(lu, y[idx], lower=True)
        x = solve_triangular(lv, a.T, lower=True)
        out[idx] = log_z - 0.5 * np.square(x).sum()
    return out
