"""Known-answer cases ported from the reference's own tests (inputs are the
reference's literals; expected values are recomputed with the same SciPy /
NumPy formulas the reference's tests use).  Shared by the CPU oracle-pin tests
and the GPU parity tests."""
import numpy as np
from scipy import stats
from scipy.special import logsumexp, gammaln


def normal_cases():
    """tests/distributions/test_univariate.py:128-152 (+ docs concepts.rst:101-105)."""
    out = []
    for given, mean, logstd in [
            (0., 0., 0.),
            ([0.99, 0.9, 9., 99.], 1., [-3., -1., 1., 10.]),
            ([7.], [0., 4.], [[1., 2.], [3., 5.]])]:
        given, mean, logstd = (np.array(a, np.float32)
                               for a in (given, mean, logstd))
        target = stats.norm.logpdf(given, mean, np.exp(logstd))
        out.append((given, mean, logstd, 0, target))
    # docs/tutorials/concepts.rst:101-105
    out.append((np.float32(0.), np.array([[-1., 1.], [0., -2.]], np.float32),
                np.float32(0.), 1,
                np.array([-2.83787704, -3.83787727])))
    return out


def bernoulli_cases():
    """tests/distributions/test_univariate.py:364-383."""
    out = []
    for logits, given in [
            (0., [0, 1]),
            ([-50., -10., -50.], [1, 1, 0]),
            ([0., 4.], [[0, 1], [0, 1]]),
            ([[2., 3., 1.], [5., 7., 4.]], np.ones([3, 1, 2, 3], np.int32))]:
        logits = np.array(logits, np.float32)
        given = np.array(given, np.float32)
        target = stats.bernoulli.logpmf(given, 1. / (1. + np.exp(-logits)))
        out.append((logits, given, target))
    return out


def categorical_cases():
    """tests/distributions/test_univariate.py:537-565."""
    out = []
    for logits, given in [
            ([0.], [0, 0, 0]),
            ([-50., -10., -50.], [0, 1, 2, 1]),
            ([0., 4.], [[0, 1], [0, 1]]),
            ([[2., 3., 1.], [5., 7., 4.]], np.ones([3, 1, 1], np.int32))]:
        logits = np.array(logits, np.float32)
        given = np.array(given, np.int32)
        nl = logits - logsumexp(logits, axis=-1, keepdims=True)
        bshape = np.broadcast_shapes(given.shape, logits.shape[:-1])
        g = np.broadcast_to(given, bshape)
        l = np.broadcast_to(nl, bshape + (logits.shape[-1],))
        target = np.take_along_axis(l, g[..., None], -1)[..., 0]
        out.append((logits, given, target))
    return out


def unnorm_multinomial_cases():
    """tests/distributions/test_multivariate.py:327-354."""
    out = []
    for normalize in [True, False]:
        for logits, given in [
                ([-50., -20., 0.], [1, 0, 3]),
                ([1., 10., 1000.], [1, 0, 0]),
                ([[2., 3., 1.], [5., 7., 4.]], np.ones([3, 1, 3], np.int32)),
                ([-10., 10., 20., 50.], [[0, 1, 99, 100], [100, 99, 1, 0]])]:
            logits = np.array(logits, np.float32)
            given = np.array(given)
            ml = logits.astype(np.float64)
            if normalize:
                ml = ml - logsumexp(ml, axis=-1, keepdims=True)
            target = np.sum(given * ml, -1)
            out.append((logits, given, normalize, target))
    return out


def _dirichlet_logpdf(x, alpha):
    lnB = np.sum(gammaln(alpha), -1) - gammaln(np.sum(alpha, -1))
    return -lnB + np.sum(np.log(x) * (alpha - 1), -1)


def dirichlet_cases():
    """tests/distributions/test_multivariate.py:509-561 (the cases whose
    target is finite; the reference's own TODO excludes alpha=1, given=0)."""
    out = []
    for alpha, given in [
            ([1., 1., 1.], [[0.2, 0.5, 0.3], [0.3, 0.4, 0.3]]),
            ([[1., 2.], [3., 4.]], [0.5, 0.5]),
            ([[5., 6.], [7., 8.]], [[0.1, 0.9]])]:
        alpha = np.array(alpha, np.float32)
        given = np.array(given, np.float32)
        gb, ab = np.broadcast_arrays(given.astype(np.float64),
                                     alpha.astype(np.float64))
        out.append((alpha, given, _dirichlet_logpdf(gb, ab)))
    return out


def mvn_params(seed, shape=(4, 5, 3)):
    """tests/distributions/test_multivariate.py:54-64 (_gen_test_params;
    batch shrunk from (10, 11) to keep the pure-Python oracle loop short)."""
    np.random.seed(seed)
    b0, b1, n = shape
    mean = 10 * np.random.normal(size=shape)
    cov = np.zeros((b0, b1, n, n))
    chol = np.zeros_like(cov)
    for i in range(b0):
        for j in range(b1):
            cov[i, j] = stats.invwishart.rvs(n, np.eye(n))
            cov[i, j] /= np.max(np.diag(cov[i, j]))
            chol[i, j] = np.linalg.cholesky(cov[i, j])
    return mean, cov, chol


LME_A = np.array([[[1., 3., 0.2], [0.7, 2., 1e-6]],
                  [[0., 1e6, 1.], [1., 1., 1.]]])       # tests/test_utils.py:259-260
LME_B = np.array([[0., 1e-6, 10.1]])                    # tests/test_utils.py:282


def univariate_more_cases():
    """The `_test_value` literals of tests/distributions/test_univariate.py for FoldNormal
    (:287-309), Uniform (:624-644), Gamma (:688-711), Beta (:764-784), Poisson (:832-858),
    Binomial (:919-950), InverseGamma (:1006-1028), Laplace (:1123-1145), BinConcrete
    (:1199-1220), each with the scipy.stats target the reference test computes.
    -> list of (family, given, a, b, target float64, atol)."""
    f32 = lambda v: np.array(v, np.float32)
    out = []
    for given, mean, logstd in [([0.99, 0.9, 9., 99.], 1., [-3., -1., 1., 10.]),
                                (0., 0., 0.),
                                ([7.], [0., 4.], [[1., 2.], [3., 5.]])]:
        g, m, ls = f32(given), f32(mean), f32(logstd)
        out.append(("fold_normal", g, m, ls,
                    stats.foldnorm.logpdf(g, m / np.exp(ls), 0, np.exp(ls)), 1e-5))
    for lo, hi, given in [(0., 1., [-1., 0., 0.5, 2.]),
                          ([-1e10, -1], [1, 1e10], 0.),
                          ([0., -1.], [[[1., 2.], [3., 5.], [4., 9.]]], [7.])]:
        lo, hi, g = f32(lo), f32(hi), f32(given)
        out.append(("uniform", g, lo, hi, stats.uniform.logpdf(g, lo, hi - lo), 1e-5))
    abg = [(1., 1., [1., 10., 1e8]),
           ([0.5, 1., 2., 3., 5., 7.5, 9.], [2., 2., 2., 1., 0.5, 1., 1.],
            np.transpose([np.arange(1, 20)])),
           ([1e-8, 1e8], [[1., 1e8], [1e-8, 5.]], [7.])]
    for a, b, given in abg:
        a, b, g = f32(a), f32(b), f32(given)
        out.append(("gamma", g, a, b, stats.gamma.logpdf(g, a, scale=1. / b), 1e-5))
        out.append(("inverse_gamma", g, a, b, stats.invgamma.logpdf(g, a, scale=b), 1e-5))
    for a, b, given in [([0.5, 5., 1., 2., 2.], [0.5, 1., 3., 2., 5.],
                         np.transpose([np.arange(0.1, 1, 0.1)])),
                        ([[1e-8], [1e8]], [[1., 1e8], [1e-8, 1.]], [0.7])]:
        a, b, g = f32(a), f32(b), f32(given)
        out.append(("beta", g, a, b, stats.beta.logpdf(g, a, b), 1e-5))
    for rate, given in [(1, [0, 1, 2, 3, 4, 5, 6]), ([5, 1, 5], [0, 0, 1]),
                        ([10000, 1], [[100, 0], [0, 100]]),
                        ([[1, 10, 100], [999, 99, 9]], np.ones([3, 1, 2, 3], np.int32))]:
        r, g = f32(rate), np.array(given, np.int32)
        out.append(("poisson", g, r, None, stats.poisson.logpmf(g, r), 1e-5))
    for logits, n, given in [(0., 6, [0, 1, 2, 3, 4, 5, 6]), ([5., -1., 5.], 2, [0, 0, 1]),
                             ([10., -10., 0.], 200, [[10, 10, 10], [190, 190, 190]]),
                             ([[1., 5., 10.], [-1., -5., -10.]], 20,
                              np.ones([3, 1, 2, 3], np.int32))]:
        l, g = np.array(logits, np.float64), np.array(given, np.int32)
        out.append(("binomial", g, f32(logits), n,
                    stats.binom.logpmf(g, n, 1 / (1. + np.exp(-l))), 1e-2))
    for loc, scale, given in [(0., 1., [.01, .1, 1., 10., 100.]),
                              ([-3, -2, -1, 0, 1, 2, 3], [.1, 3, 2, 3, 3, 2, .1],
                               np.transpose([np.arange(1, 20)])),
                              ([1e-5, -1e-5], [[1., 10.], [1e8, 5.]], [7.])]:
        m, s, g = f32(loc), f32(scale), f32(given)
        out.append(("laplace", g, m, s, stats.laplace.logpdf(g, m, scale=s), 1e-5))
    for given, t, logits in [([0.001, 0.01, 0.1, 0.5, 0.9, 0.99, 0.999], tl[0], tl[1])
                             for tl in [(0.1, 0.1), (0.01, 0.5), (0.66, 0.9), (1., 0.99)]]:
        g = np.array(given, np.float64)
        tgt = np.log(t) + logits - (t + 1) * np.log(g) - (t + 1) * np.log(1 - g) - \
            2 * np.log(np.exp(logits) * (g ** -t) + (1 - g) ** -t)
        out.append(("bin_concrete", f32(given), f32(t), f32(logits), tgt, 1e-4))
    return out


def multinomial_cases():
    """tests/distributions/test_multivariate.py:218-253 `_test_value` literals with the target the
    reference test computes (factorials + normalised logits).
    -> list of (logits, n_experiments, given, normalize_logits, target)."""
    from scipy.special import factorial, logsumexp
    out = []
    for normalize in (True, False):
        for logits, n, given in [([-50., -20., 0.], 4, [1, 0, 3]),
                                 ([1., 10., 1000.], 1, [1, 0, 0]),
                                 ([[2., 3., 1.], [5., 7., 4.]], 7, np.array([3, 1, 3], np.int32)),
                                 ([-10., 10., 20., 50.], 100, [[0, 1, 49, 50], [50, 49, 1, 0]])]:
            l = np.array(logits, np.float32)
            g = np.array(given)
            ml = l - logsumexp(l, axis=-1, keepdims=True) if normalize else l
            ne = np.sum(g, axis=-1)
            tgt = np.log(factorial(ne)) - np.sum(np.log(factorial(g)), -1) + np.sum(g * ml, -1)
            out.append((l, n, g, normalize, tgt))
    return out
