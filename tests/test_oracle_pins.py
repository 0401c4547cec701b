"""Pins the CPU oracle to the reference's own known-answer tests and golden
vectors.  CPU only."""
import numpy as np
import pytest
from scipy import stats
from scipy.special import logsumexp

import cases
from oracle import distributions as OD
from oracle import variational as OV
from oracle import philox


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-6), (np.float64, 1e-12)])
def test_normal_log_prob(dtype, tol):
    for given, mean, logstd, g, target in cases.normal_cases():
        lp = OD.normal_log_prob(given, mean, logstd, g, dtype)
        np.testing.assert_allclose(lp, target, rtol=max(tol, 1e-6), atol=1e-6)


def test_bernoulli_log_prob():
    for logits, given, target in cases.bernoulli_cases():
        np.testing.assert_allclose(OD.bernoulli_log_prob(given, logits),
                                   target, rtol=1e-6, atol=1e-6)


def test_categorical_log_prob():
    for logits, given, target in cases.categorical_cases():
        np.testing.assert_allclose(OD.categorical_log_prob(given, logits),
                                   target, rtol=1e-6, atol=1e-6)


def test_unnormalized_multinomial_log_prob():
    for logits, given, norm, target in cases.unnorm_multinomial_cases():
        lp = OD.unnormalized_multinomial_log_prob(given, logits, norm)
        np.testing.assert_allclose(lp, target, rtol=1e-5, atol=1e-4)


def test_dirichlet_log_prob():
    for alpha, given, target in cases.dirichlet_cases():
        np.testing.assert_allclose(OD.dirichlet_log_prob(given, alpha),
                                   target, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("seed", [23, 233, 2333])
def test_mvn_cholesky_log_prob(seed):
    mean, cov, chol = cases.mvn_params(seed)
    rng = np.random.RandomState(seed)
    samples = mean + np.einsum('ijab,nijb->nija', chol,
                               rng.standard_normal((7,) + mean.shape))
    lp = OD.mvn_cholesky_log_prob(samples, mean, chol, dtype=np.float64)
    for i in range(mean.shape[0]):
        for j in range(mean.shape[1]):
            exact = stats.multivariate_normal.logpdf(
                samples[:, i, j, :], mean[i, j], cov[i, j])
            np.testing.assert_allclose(lp[:, i, j], exact, rtol=1e-8,
                                       atol=1e-8)


def test_log_mean_exp_golden():
    """tests/test_utils.py:270-284."""
    a = cases.LME_A
    for keepdims in [True, False]:
        true = logsumexp(a, (0, 2), keepdims=keepdims) - np.log(
            a.shape[0] * a.shape[2])
        np.testing.assert_allclose(
            OV.log_mean_exp(a, (0, 2), keepdims, np.float64), true,
            rtol=1e-6)
        true_s = logsumexp(a, (0, 2), keepdims=keepdims)
        np.testing.assert_allclose(
            OV.log_sum_exp(a, (0, 2), keepdims, np.float64), true_s,
            rtol=1e-6)
    b = cases.LME_B
    assert np.abs(OV.log_mean_exp(b, 0, False, np.float64) - b).max() < 1e-6


def _kl_normal_normal(m1, s1, m2, s2):
    return np.log(s2 / s1) + (s1 ** 2 + (m1 - m2) ** 2) / (2 * s2 ** 2) - 0.5


def _n01(shape):
    return np.random.RandomState(1).standard_normal(shape).astype(np.float32)


@pytest.mark.parametrize("x_mean,x_std", [(0., 1.), (2., 3.)])
def test_elbo_value_vs_analytic_kl(x_mean, x_std):
    """tests/variational/test_exclusive_kl.py:26-47: q-samples are
    RandomState(1).standard_normal(1e5), p = N(x_mean, x_std);
    ELBO == -KL(N(0,1) || p) to 1e-3."""
    z = _n01(100000)
    log_q = stats.norm.logpdf(z).astype(np.float32)
    log_p = OD.normal_log_prob(z, x_mean, np.log(x_std))
    lb = OV.elbo(log_p, [log_q], axis=0)
    assert abs(lb - (-_kl_normal_normal(0., 1., x_mean, x_std))) < 1e-3


def _sgvb_grads(weights, z, eps, x_mean, x_std, sigma):
    """d(-objective)/d(mu, sigma) through z = mu + sigma*eps with backward
    weights d objective / d log_w (1/K for ELBO, softmax for IWAE):
    log_w = log p(z) - log q(z); dlog p/dz = -(z-m)/s^2;
    log q(z(mu,sigma)) = -log sigma - eps^2/2 + c."""
    dlogp_dz = -(z - x_mean) / x_std ** 2
    dmu = -(weights * dlogp_dz).sum()
    dsigma = -(weights * (dlogp_dz * eps + 1.0 / sigma)).sum()
    return dmu, dsigma


@pytest.mark.parametrize("x_mean,x_std,rtol,atol",
                         [(0., 1., 1e-2, 1e-6), (2., 3., 1e-6, 1e-2)])
def test_elbo_sgvb_gradient_vs_kl_grads(x_mean, x_std, rtol, atol):
    """tests/variational/test_exclusive_kl.py:49-78 (q = N(2, 3))."""
    mu, sigma = 2., 3.
    eps = _n01(100000).astype(np.float64)
    z = eps * sigma + mu
    w = OV.elbo_grad_logw(z.shape, 0, np.float64)
    g = _sgvb_grads(w, z, eps, x_mean, x_std, sigma)
    true = ((mu - x_mean) / x_std ** 2, -1.0 / sigma + sigma / x_std ** 2)
    np.testing.assert_allclose(g, true, rtol=rtol, atol=atol)


@pytest.mark.parametrize("x_mean,x_std", [(0., 1.), (2., 3.)])
def test_iwae_k1_equals_elbo_and_monotone(x_mean, x_std):
    """tests/variational/test_monte_carlo.py:25-70."""
    rng = np.random.RandomState(1)
    n1 = rng.standard_normal(size=(1, 1000)).astype(np.float32)
    n3 = rng.standard_normal(1000).astype(np.float32)
    analytic = -_kl_normal_normal(0., 1., x_mean, x_std)
    log_w = OD.normal_log_prob(n1, x_mean, np.log(x_std)) - \
        stats.norm.logpdf(n1).astype(np.float32)
    k1 = OV.iw_objective(log_w, [], axis=0).mean()
    assert abs(k1 - analytic) < 1e-2
    log_w3 = OD.normal_log_prob(n3, x_mean, np.log(x_std)) - \
        stats.norm.logpdf(n3).astype(np.float32)
    k1000 = OV.iw_objective(log_w3, [], axis=0).mean()
    assert k1000 > analytic - 1e-6
    with pytest.raises(ValueError):
        OV.iw_objective(log_w, [], axis=None)


@pytest.mark.parametrize("x_mean,x_std,thr", [(0., 1., 0.04), (2., 3., 0.02)])
def test_iwae_sgvb_gradient_vs_kl_grads(x_mean, x_std, thr):
    """tests/variational/test_monte_carlo.py:72-102 (K=1 along axis 0)."""
    mu, sigma = 2., 3.
    rng = np.random.RandomState(1)
    eps = rng.standard_normal(size=(1, 1000)).astype(np.float32).astype(
        np.float64)
    z = eps * sigma + mu
    log_w = stats.norm.logpdf(z, x_mean, x_std) - stats.norm.logpdf(
        z, mu, sigma)
    w = OV.iw_grad_logw(log_w, 0, np.float64) / log_w.shape[1]  # reduce_mean
    g = _sgvb_grads(w, z, eps, x_mean, x_std, sigma)
    true = ((mu - x_mean) / x_std ** 2, -1.0 / sigma + sigma / x_std ** 2)
    np.testing.assert_allclose(g, true, rtol=thr, atol=thr)


def test_iw_grad_is_softmax():
    x = np.random.RandomState(0).standard_normal((6, 5))
    w = OV.iw_grad_logw(x, 0, np.float64)
    np.testing.assert_allclose(w.sum(0), 1.0, rtol=1e-12)
    h = 1e-6
    xp = x.copy()
    xp[2, 3] += h
    fd = (OV.log_mean_exp(xp, 0, dtype=np.float64).sum()
          - OV.log_mean_exp(x, 0, dtype=np.float64).sum()) / h
    np.testing.assert_allclose(fd, w[2, 3], rtol=1e-4)


def test_philox_known_answers():
    """Random123 philox4x32-10 KAT vectors."""
    def h(a):
        return [int(v) for v in a]
    z4, z2 = np.zeros(4, np.uint32), np.zeros(2, np.uint32)
    assert h(philox.philox4x32_10(z4, z2)) == [
        0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f4 = np.full(4, 0xffffffff, np.uint32)
    f2 = np.full(2, 0xffffffff, np.uint32)
    assert h(philox.philox4x32_10(f4, f2)) == [
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    c = np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], np.uint32)
    k = np.array([0xa4093822, 0x299f31d0], np.uint32)
    assert h(philox.philox4x32_10(c, k)) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normals_are_standard():
    z = philox.normal_matrix(seed=7, stream=1, iteration=3, row0=0,
                             n_rows=2000, n_cols=16)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    u = philox.uniform_vector(7, 2, 3, 0, 20000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01


def _dlogq(x, mu, sigma):
    """d log N(x; mu, sigma) / d(mu, sigma) with the sample x held fixed."""
    return (x - mu) / sigma ** 2, -1.0 / sigma + (x - mu) ** 2 / sigma ** 3


@pytest.mark.parametrize("x_mean,x_std,thr", [(0., 1., 1e-2), (2., 3., 1e-6)])
def test_vimco_gradient_matches_sgvb(x_mean, x_std, thr):
    """tests/variational/test_monte_carlo.py:104-142: with q = N(2, 3) and the
    1000 `_n3_samples`, the VIMCO gradient wrt (mu, sigma) equals the SGVB one
    to the reference's thresholds."""
    mu, sigma = 2., 3.
    rng = np.random.RandomState(1)
    rng.standard_normal(size=(1, 1000))           # _n1_samples come first
    eps = rng.standard_normal(1000).astype(np.float32).astype(np.float64)
    x = eps * sigma + mu
    log_w = stats.norm.logpdf(x, x_mean, x_std) - stats.norm.logpdf(x, mu, sigma)
    # SGVB (reparameterised): -d LME / d(mu, sigma)
    g_sgvb = _sgvb_grads(OV.iw_grad_logw(log_w, 0, np.float64), x, eps,
                         x_mean, x_std, sigma)
    # VIMCO (score function): samples fixed, gradient through log q only
    gq = OV.vimco_grad_logq(log_w, 0, np.float64)
    dmu, dsig = _dlogq(x, mu, sigma)
    g_vimco = ((gq * dmu).sum(), (gq * dsig).sum())
    np.testing.assert_allclose(g_vimco, g_sgvb, rtol=thr, atol=thr)
    with pytest.raises(ValueError, match="larger than 1"):
        OV.vimco_signal(log_w[:1], 0)


def test_vimco_signal_is_leave_one_out():
    """Direct definition: signal[k] = LME(l) - LME(l with l_k := mean of others)."""
    l = np.random.RandomState(3).standard_normal((5, 7)) * 3
    sig = OV.vimco_signal(l, 0, np.float64)
    for k in range(5):
        for i in range(7):
            col = l[:, i].copy()
            col[k] = (l[:, i].sum() - l[k, i]) / 4
            want = OV.log_mean_exp(l[:, i], 0, dtype=np.float64) - \
                OV.log_mean_exp(col, 0, dtype=np.float64)
            assert abs(sig[k, i] - want) < 1e-12
    sig1 = OV.vimco_signal(l.T, 1, np.float64)        # other axis
    np.testing.assert_allclose(sig1, sig.T, rtol=1e-12)


@pytest.mark.parametrize("x_mean,x_std,thr", [(0., 1., 0.01), (2., 3., 0.02)])
def test_importance_gradient_vs_kl_grads(x_mean, x_std, thr):
    """tests/variational/test_inclusive_kl.py:44-72: self-normalised importance
    gradient of KL(p || q) wrt (mu, sigma), q = N(2, 3), 1e5 samples."""
    mu, sigma = 2., 3.
    eps = _n01(100000).astype(np.float64)
    x = eps * sigma + mu
    log_q = stats.norm.logpdf(x, mu, sigma)
    w = OV.normalized_weights(stats.norm.logpdf(x, x_mean, x_std) - log_q, 0,
                              np.float64)
    dmu, dsig = _dlogq(x, mu, sigma)
    g = (-(w * dmu).sum(), -(w * dsig).sum())       # cost = sum w~ * (-log q)
    # d KL(p || q) / d(mu, sigma), p = N(x_mean, x_std)
    true = (-(x_mean - mu) / sigma ** 2,
            1.0 / sigma - (x_std ** 2 + (x_mean - mu) ** 2) / sigma ** 3)
    np.testing.assert_allclose(g, true, rtol=thr, atol=thr)
    c = OV.importance_cost(stats.norm.logpdf(x, x_mean, x_std), log_q, 0, np.float64)
    assert np.isfinite(c)


@pytest.mark.parametrize("x_mean,x_std,rtol,atol", [(0., 1., 1e-2, 1e-6), (2., 3., 1e-6, 1e-6)])
def test_reinforce_gradient_vs_kl_grads(x_mean, x_std, rtol, atol):
    """tests/variational/test_exclusive_kl.py:80-112 (variance_reduction=False,
    q = N(2, 3), the 1e6-sample stream that follows the 1e5 one)."""
    mu, sigma = 2., 3.
    rng = np.random.RandomState(1)
    rng.standard_normal(100000)
    eps = rng.standard_normal(1000000).astype(np.float32).astype(np.float64)
    x = eps * sigma + mu
    lj, lq = stats.norm.logpdf(x, x_mean, x_std), stats.norm.logpdf(x, mu, sigma)
    gq = OV.reinforce_grad_logq(lj, lq, 0, np.float64)
    dmu, dsig = _dlogq(x, mu, sigma)
    g = ((gq * dmu).sum(), (gq * dsig).sum())
    true = ((mu - x_mean) / x_std ** 2, -1.0 / sigma + sigma / x_std ** 2)
    np.testing.assert_allclose(g, true, rtol=rtol, atol=atol)


def test_univariate_more_log_prob_vs_reference_targets():
    """Oracle restatements of the nine other univariate densities against the scipy.stats
    targets of the reference's own `_test_value` cases (tests/distributions/test_univariate.py)."""
    n = 0
    for fam, given, a, b, target, atol in cases.univariate_more_cases():
        fn = getattr(OD, fam + "_log_prob")
        with np.errstate(all="ignore"):
            got = fn(given, a, b, dtype=np.float64) if b is not None else \
                fn(given, a, dtype=np.float64)
        np.testing.assert_allclose(got, target, rtol=1e-6, atol=atol, err_msg=fam)
        n += 1
    assert n == 29


def test_multinomial_and_onehot_vs_reference_targets():
    """tests/distributions/test_multivariate.py:218-253 (Multinomial) and 505-533
    (OnehotCategorical: one-hot of the Categorical cases, target = normalised logit)."""
    for l, n, g, normalize, tgt in cases.multinomial_cases():
        for ne in (None, n):
            got = OD.multinomial_log_prob(g, l, ne, normalize, dtype=np.float64)
            np.testing.assert_allclose(got, tgt, rtol=1e-6, atol=1e-6)
    logits = np.array([[2., 3., 1.], [5., 7., 4.]], np.float32)
    idx = np.array([1, 0])
    onehot = np.eye(3)[idx]
    want = (logits - np.log(np.exp(logits).sum(-1, keepdims=True)))[np.arange(2), idx]
    np.testing.assert_allclose(OD.onehot_categorical_log_prob(onehot, logits, dtype=np.float64),
                               want, rtol=1e-6)
    np.testing.assert_allclose(OD.onehot_categorical_log_prob(onehot, logits, dtype=np.float64),
                               OD.categorical_log_prob(idx, logits, dtype=np.float64), rtol=1e-6)


def test_concrete_family_and_matrix_normal_vs_reference_targets():
    """tests/distributions/test_multivariate.py: ExpConcrete :735-759 and Concrete :870-894
    `_test_value` literals (targets restated as the tests compute them); the matrix-variate
    normal against scipy.stats.matrix_normal as in :1030-1046."""
    from scipy.special import gammaln
    for given, t, logits in [([0.25, 0.25, 0.5], 0.1, [1., 1., 1.2]),
                             ([[0.25, 0.25, 0.5], [0.1, 0.5, 0.4]], 0.5,
                              [[1., 1., 1.], [.5, .5, .4]])]:
        g = np.array(given, np.float32).astype(np.float64)
        l = np.array(logits, np.float32).astype(np.float64)
        n = l.shape[-1]
        tgt = gammaln(n) + (n - 1) * np.log(t) + (l - t * np.log(g) - np.log(g)).sum(-1) - \
            n * np.log(np.exp(l - t * np.log(g)).sum(-1))
        np.testing.assert_allclose(OD.concrete_log_prob(g, t, l, dtype=np.float64), tgt, rtol=1e-6)
        lg = np.log(g)
        tgt_e = gammaln(n) + (n - 1) * np.log(t) + (l - t * lg).sum(-1) - \
            n * np.log(np.exp(l - t * lg).sum(-1))
        np.testing.assert_allclose(OD.exp_concrete_log_prob(lg, t, l, dtype=np.float64), tgt_e,
                                   rtol=1e-6)
    rng = np.random.RandomState(23)
    r, c = 3, 4
    a = rng.standard_normal((r, r)); u = a @ a.T + r * np.eye(r)
    b = rng.standard_normal((c, c)); v = b @ b.T + c * np.eye(c)
    mean = rng.standard_normal((r, c))
    x = rng.standard_normal((5, r, c))
    got = OD.matrix_normal_cholesky_log_prob(x, mean, np.linalg.cholesky(u), np.linalg.cholesky(v))
    np.testing.assert_allclose(got, stats.matrix_normal.logpdf(x, mean, u, v), rtol=1e-9)
