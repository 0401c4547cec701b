"""GPU parity / moment tests of the device samplers (csrc/samplers.cu) behind
Categorical / OnehotCategorical / Multinomial / Dirichlet / Gamma-family `.sample()`
(zhusuan/distributions/univariate.py:478-494, multivariate.py:660-663)."""
import numpy as np
import pytest
import torch

from oracle import samplers as OS

pytestmark = pytest.mark.gpu


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def zs():
    import zhusuan_b200 as zs
    return zs


@pytest.mark.parametrize("C", [2, 7, 33, 1000])
def test_categorical_injected_uniforms_match_oracle(zs, C):
    rng = np.random.RandomState(C)
    logits = (2.0 * rng.standard_normal((5, 3, C))).astype(np.float32)
    logits[0, 0, : C // 2] = -np.inf                      # zero-mass categories are never drawn
    u = rng.random_sample((11, 5, 3)).astype(np.float32)
    u[0, 0, 0], u[1, 0, 0] = 0.0, np.float32(1.0) - np.float32(2.0 ** -24)
    d = zs.distributions.Categorical(T(logits))
    out = N(d._sample(11, u=T(u)))
    assert out.dtype == np.int32 and out.shape == (11, 5, 3)
    ref = OS.categorical_inverse_cdf(logits.reshape(-1, C), u.reshape(11, -1)).reshape(out.shape)
    margin = OS.cdf_margin(logits.reshape(-1, C), u.reshape(11, -1)).reshape(out.shape)
    bad = out != ref
    assert np.all(margin[bad] < 1e-5), "mismatch away from a CDF step"
    assert bad.mean() < 0.01
    assert np.all(out[:, 0, 0] >= C // 2)


def test_categorical_philox_stream_and_frequencies(zs):
    rng = np.random.RandomState(1)
    logits = rng.standard_normal((4, 6)).astype(np.float32)
    zs.set_random_seed(77)
    d = zs.distributions.Categorical(T(logits))
    seed_it = []
    orig = d._next_rng
    d._next_rng = lambda: seed_it.append(orig()) or seed_it[-1]
    n = 20000
    out = N(d.sample(n))
    seed, it = seed_it[0]
    u = OS.categorical_uniforms(seed, it, n * 4).reshape(n, 4)
    ref = OS.categorical_inverse_cdf(logits, u)
    margin = OS.cdf_margin(logits, u)
    bad = out != ref
    assert np.all(margin[bad] < 1e-5) and bad.mean() < 1e-3
    p = np.exp(logits - logits.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    freq = np.stack([(out == c).mean(0) for c in range(6)], -1)
    np.testing.assert_allclose(freq, p, atol=4 * np.sqrt(0.25 / n))


def test_onehot_and_multinomial_use_the_device_sampler(zs):
    logits = T(np.log(np.array([[0.2, 0.3, 0.5], [0.6, 0.3, 0.1]], np.float32)))
    oh = zs.distributions.OnehotCategorical(logits).sample(4000)
    assert tuple(oh.shape) == (4000, 2, 3) and oh.dtype == torch.int32
    assert torch.all(oh.sum(-1) == 1)
    np.testing.assert_allclose(N(oh.float().mean(0)), [[0.2, 0.3, 0.5], [0.6, 0.3, 0.1]],
                               atol=0.04)
    m = zs.distributions.Multinomial(logits, n_experiments=10).sample(3000)
    assert tuple(m.shape) == (3000, 2, 3) and torch.all(m.sum(-1) == 10)
    np.testing.assert_allclose(N(m.float().mean(0)), [[2, 3, 5], [6, 3, 1]], atol=0.15)


def test_dirichlet_philox_matches_oracle_and_moments(zs):
    rng = np.random.RandomState(3)
    alpha = (0.3 + 4 * rng.random_sample((3, 5))).astype(np.float32)
    alpha[0, 0] = 0.2                                         # alpha < 1: boosted branch
    zs.set_random_seed(5)
    d = zs.distributions.Dirichlet(T(alpha))
    seed_it = []
    orig = d._next_rng
    d._next_rng = lambda: seed_it.append(orig()) or seed_it[-1]
    n = 4000
    x = N(d.sample(n))
    assert x.shape == (n, 3, 5)
    np.testing.assert_allclose(x.sum(-1), 1.0, rtol=1e-5)
    seed, it = seed_it[0]
    ref = OS.dirichlet(alpha, n, seed, it)
    close = np.abs(x - ref) <= 2e-4 * np.maximum(ref, 1e-3)
    assert close.all(-1).mean() > 0.995      # a float32 squeeze test may flip at its boundary
    mean = alpha / alpha.sum(-1, keepdims=True)
    a0 = alpha.sum(-1, keepdims=True)
    var = mean * (1 - mean) / (a0 + 1)
    np.testing.assert_allclose(x.mean(0), mean, atol=5 * np.sqrt(var.max() / n))
    np.testing.assert_allclose(x.var(0), var, rtol=0.25, atol=2e-4)


def test_dirichlet_injected_gammas_normalise(zs):
    rng = np.random.RandomState(4)
    alpha = (0.5 + rng.random_sample((2, 4))).astype(np.float32)
    g = rng.gamma(np.broadcast_to(alpha, (6, 2, 4))).astype(np.float32)
    x = N(zs.distributions.Dirichlet(T(alpha))._sample(6, gammas=T(g)))
    np.testing.assert_allclose(x, g / g.sum(-1, keepdims=True), rtol=2e-6)


def test_gamma_family_moments(zs):
    alpha = np.array([0.4, 1.0, 2.5, 9.0], np.float32)
    beta = np.array([1.0, 2.0, 0.5, 3.0], np.float32)
    n = 40000
    x = N(zs.distributions.Gamma(T(alpha), T(beta)).sample(n))
    assert x.shape == (n, 4) and (x > 0).all()
    np.testing.assert_allclose(x.mean(0), alpha / beta, rtol=0.03)
    np.testing.assert_allclose(x.var(0), alpha / beta ** 2, rtol=0.08)
    y = N(zs.distributions.Beta(T(alpha), T(beta)).sample(n))
    np.testing.assert_allclose(y.mean(0), alpha / (alpha + beta), rtol=0.03, atol=2e-3)


def test_gamma_family_sample_shapes(zs):
    """The broadcast / n_samples shape cases of the reference's test_sample_shape_2parameter
    (tests/distributions/utils.py) for the distributions that draw on the device sampler."""
    D = zs.distributions
    cases = [([2, 3], [], None, [2, 3]), ([2, 3], [], 1, [1, 2, 3]), ([5], [5], 2, [2, 5]),
             ([2, 1, 4], [1, 2, 4], 3, [3, 2, 2, 4]), ([2, 3], [2, 1], 1, [1, 2, 3]),
             ([1, 3], [], 2, [2, 1, 3]), ([2, 1, 5], [3, 1], 3, [3, 2, 3, 5])]
    for make in (D.Gamma, D.Beta, D.InverseGamma,
                 lambda a, b: D.FoldNormal(a, std=b), lambda a, b: D.Uniform(a, a + b),
                 D.Laplace):
        for s1, s2, n, target in cases:
            x = make(torch.ones(s1, device="cuda") + 1, torch.ones(s2, device="cuda") + 1).sample(n)
            assert list(x.shape) == target and x.dtype == torch.float32
            assert bool(torch.isfinite(x).all())


def test_vector_valued_sample_shapes(zs):
    """test_sample_shape_1parameter(is_univariate=False) cases for the distributions whose draws
    come from the device samplers."""
    D = zs.distributions
    makes = {"OnehotCategorical": lambda l: D.OnehotCategorical(l),
             "Multinomial": lambda l: D.Multinomial(l, 7),
             "Dirichlet": lambda l: D.Dirichlet(l.abs() + 1),
             "Concrete": lambda l: D.Concrete(torch.tensor(1., device="cuda"), l),
             "ExpConcrete": lambda l: D.ExpConcrete(torch.tensor(1., device="cuda"), l),
             "Categorical": lambda l: D.Categorical(l)}
    for name, make in makes.items():
        for shape, n, target in (([2, 4], None, [2, 4]), ([3], 2, [2, 3]),
                                 ([2, 1, 4], 3, [3, 2, 1, 4])):
            x = make(torch.zeros(shape, device="cuda")).sample(n)
            want = target[:-1] if name == "Categorical" else target
            assert list(x.shape) == want, (name, shape, n, tuple(x.shape))
    mv = D.MatrixVariateNormalCholesky(torch.zeros(5, 2, 3, device="cuda"),
                                       torch.eye(2, device="cuda").expand(5, 2, 2),
                                       torch.eye(3, device="cuda").expand(5, 3, 3))
    assert list(mv.sample(4).shape) == [4, 5, 2, 3] and list(mv.sample().shape) == [5, 2, 3]
    b = D.BinConcrete(torch.tensor(0.5, device="cuda"), torch.zeros(2, 3, device="cuda"))
    assert list(b.sample(4).shape) == [4, 2, 3]


def test_poisson_and_binomial_device_samplers(zs):
    """Poisson / Binomial draws by inverse transform from the mode (csrc/samplers.cu
    count_sample_kernel): injected uniforms vs the float64 oracle restatement (identical away from
    CDF steps), Philox stream position, exact moments incl. a large rate and p near 0 / 1."""
    D = zs.distributions
    rng = np.random.RandomState(8)
    rate = np.array([0.05, 0.7, 3.0, 12.5, 140.0], np.float32)
    u = rng.random_sample((40, 5)).astype(np.float32)
    got = N(D.Poisson(T(rate))._sample(40, u=T(u)))
    ref = OS.poisson_inverse(rate, u)
    assert got.dtype == np.int32 and (got != ref).mean() < 0.02
    logits = np.array([-4.0, -0.5, 0.0, 1.5, 6.0], np.float32)
    gotb = N(D.Binomial(T(logits), 37)._sample(40, u=T(u)))
    refb = OS.binomial_inverse(logits, 37, u)
    assert (gotb != refb).mean() < 0.02 and gotb.min() >= 0 and gotb.max() <= 37
    # Philox mode: the uniforms are words of blocks (i // 4, 0, it, STREAM_COUNT)
    zs.set_random_seed(31)
    d = D.Poisson(T(rate))
    seed_it = []
    orig = d._next_rng
    d._next_rng = lambda: seed_it.append(orig()) or seed_it[-1]
    n = 20000
    x = N(d.sample(n))
    seed, it = seed_it[0]
    uu = OS.count_uniforms(seed, it, n * 5).reshape(n, 5)
    assert (x[:200] != OS.poisson_inverse(rate, uu[:200])).mean() < 0.02
    np.testing.assert_allclose(x.mean(0), rate, rtol=0.05, atol=0.01)
    np.testing.assert_allclose(x.var(0), rate, rtol=0.08, atol=0.01)
    y = N(D.Binomial(T(logits), 37).sample(n)).astype(np.float64)
    p = 1 / (1 + np.exp(-logits.astype(np.float64)))
    np.testing.assert_allclose(y.mean(0), 37 * p, rtol=0.03, atol=0.02)
    np.testing.assert_allclose(y.var(0), 37 * p * (1 - p), rtol=0.08, atol=0.02)


def test_count_sample_shapes(zs):
    D = zs.distributions
    for make in (lambda p: D.Poisson(p.abs() + 1), lambda p: D.Binomial(p, 10)):
        for shape, n, target in (([2, 3], None, [2, 3]), ([5], 2, [2, 5]), ([1, 3], 1, [1, 1, 3])):
            x = make(torch.zeros(shape, device="cuda")).sample(n)
            assert list(x.shape) == target and x.dtype == torch.int32
