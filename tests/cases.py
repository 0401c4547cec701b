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
