"""NumPy restatement of the device samplers (TEST ORACLE ONLY -- see oracle/__init__).

The reference draws these with TensorFlow ops whose streams cannot be reproduced without TF
(``tf.random.categorical`` univariate.py:478-494, ``tf.random_gamma`` multivariate.py:660-663);
the product fixes an algorithm on its Philox generator (zhusuan_b200/csrc/samplers.cu) and this
file restates it: inverse-CDF categorical draws and Marsaglia-Tsang gamma variates.  "Parity
unpinned" by the reference (it tests shapes and moments only, tests/distributions/utils.py); the
moments are checked against the exact distribution moments in tests/test_gpu_samplers.py.
"""
import numpy as np

from . import philox as PH

STREAM_CATEGORICAL = 6
STREAM_GAMMA = 7


def categorical_uniforms(seed, it, n_draws):
    """The uniform of draw d: word 0 of Philox block (0, d, it, STREAM_CATEGORICAL)."""
    d = np.arange(n_draws, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    ctr = PH.counter(STREAM_CATEGORICAL, it, d.astype(np.uint32), np.uint32(0))
    return PH.u32_to_uniform(PH.philox4x32_10(ctr, key)[..., 0])


def categorical_inverse_cdf(logits, u):
    """logits [rows, C], u [n_samples, rows] -> int32 [n_samples, rows]: the first category whose
    cumulative softmax mass exceeds u * total (float64 here; the kernel sums in float32, so
    a draw may differ where u * total lands within rounding of a CDF step: see `cdf_margin`)."""
    l = np.asarray(logits, np.float64)
    e = np.exp(l - l.max(-1, keepdims=True))
    cdf = np.cumsum(e, -1)
    tgt = np.asarray(u, np.float64) * cdf[:, -1]
    idx = (cdf[None] > tgt[..., None]).argmax(-1)
    none = ~(cdf[None] > tgt[..., None]).any(-1)
    last = (e > 0).shape[-1] - 1 - (e > 0)[:, ::-1].argmax(-1)
    idx = np.where(none, last[None], idx)
    return idx.astype(np.int32)


def cdf_margin(logits, u):
    """min_c |u - cdf_c / total|: draws with a margin below ~1e-5 are rounding-ambiguous."""
    l = np.asarray(logits, np.float64)
    e = np.exp(l - l.max(-1, keepdims=True))
    cdf = np.cumsum(e, -1) / e.sum(-1, keepdims=True)
    return np.abs(cdf[None] - np.asarray(u, np.float64)[..., None]).min(-1)


def gamma_marsaglia_tsang(alpha, seed, it, max_attempts=64):
    """Gamma(alpha, 1) for a flat float array `alpha` (element index e = position), attempt k of
    element e using Philox block (k, e, it, STREAM_GAMMA): words (x, y) -> Box-Muller normal,
    z -> uniform (0, 1], w of attempt 0 -> the alpha < 1 boost u^(1/alpha)."""
    a0 = np.asarray(alpha, np.float64).reshape(-1)
    n = a0.size
    boost = a0 < 1.0
    a = np.where(boost, a0 + 1.0, a0)
    d = a - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    g = d.copy()
    done = np.zeros(n, bool)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    e = np.arange(n, dtype=np.uint32)
    w_boost = None
    for k in range(max_attempts):
        w = PH.philox4x32_10(PH.counter(STREAM_GAMMA, it, e, np.uint32(k)), key)
        if k == 0:
            w_boost = w[..., 3]
        z0, _ = PH.box_muller(w[..., 0], w[..., 1])
        z0 = z0.astype(np.float64)
        v1 = 1.0 + c * z0
        ok = v1 > 0
        v = np.where(ok, v1, 1.0) ** 3
        u = PH.u32_to_uniform_open(w[..., 2]).astype(np.float64)
        acc = ok & (np.log(u) < 0.5 * z0 * z0 + d - d * v + d * np.log(v))
        take = acc & ~done
        g[take] = (d * v)[take]
        done |= acc
        if done.all():
            break
    ub = PH.u32_to_uniform_open(w_boost).astype(np.float64)
    g = np.where(boost, g * ub ** (1.0 / a0), g)
    return g.reshape(np.shape(alpha))


def dirichlet(alpha, n_samples, seed, it):
    """[n_samples] + alpha.shape: row r of the flattened [n_samples * rows, C] output uses
    alpha row r % rows and element index e = r * C + c."""
    a = np.asarray(alpha, np.float64)
    C = a.shape[-1]
    full = np.broadcast_to(a, (n_samples,) + a.shape).reshape(-1)
    g = gamma_marsaglia_tsang(full, seed, it).reshape(-1, C)
    return (g / g.sum(-1, keepdims=True)).reshape((n_samples,) + a.shape)


STREAM_COUNT = 8


def count_uniforms(seed, it, n):
    """The uniform of draw i: word i % 4 of Philox block (i // 4, 0, it, STREAM_COUNT)."""
    i = np.arange(n, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    ctr = PH.counter(STREAM_COUNT, it, np.uint32(0), (i >> np.uint64(2)).astype(np.uint32))
    w = PH.philox4x32_10(ctr, key)
    return PH.u32_to_uniform(w[np.arange(n), (i & np.uint64(3)).astype(np.int64)])


def _invert_from_mode(u, mode, pmf, kmax):
    """Inverse transform visiting mode, mode+1, mode-1, mode+2, ... (float64 pmf callable)."""
    s = pmf(mode)
    if u < s:
        return mode
    lo = hi = mode
    for _ in range(10 * (kmax if kmax < 10 ** 8 else 10 ** 5) + 100):
        if hi < kmax:
            hi += 1
            s += pmf(hi)
            if u < s:
                return hi
        if lo > 0:
            lo -= 1
            s += pmf(lo)
            if u < s:
                return lo
        if hi >= kmax and lo <= 0:
            break
    return hi


def poisson_inverse(rate, u):
    """Poisson draws from uniforms, the kernel's enumeration order (float64 pmf)."""
    from scipy import stats
    rate, u = np.broadcast_arrays(np.asarray(rate, np.float64), np.asarray(u, np.float64))
    out = np.empty(u.shape, np.int32)
    for idx in np.ndindex(u.shape):
        lam = rate[idx]
        out[idx] = _invert_from_mode(u[idx], int(np.floor(lam)),
                                     lambda k: stats.poisson.pmf(k, lam), 10 ** 9) if lam > 0 else 0
    return out


def binomial_inverse(logits, n, u):
    from scipy import stats
    logits, u = np.broadcast_arrays(np.asarray(logits, np.float64), np.asarray(u, np.float64))
    out = np.empty(u.shape, np.int32)
    for idx in np.ndindex(u.shape):
        p = 1.0 / (1.0 + np.exp(-logits[idx]))
        m = min(n, int(np.floor((n + 1) * p)))
        out[idx] = _invert_from_mode(u[idx], m, lambda k: stats.binom.pmf(k, n, p), n)
    return out
