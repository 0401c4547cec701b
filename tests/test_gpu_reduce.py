"""GPU parity: K6 reductions (log_mean_exp / mean and their backward) and the
ELBO / IWAE objectives, against the reference's golden array
(tests/test_utils.py:257-284), the seeded analytic-KL tests
(tests/variational/*.py) and the CPU oracle."""
import numpy as np
import pytest
import torch
from scipy import stats
from scipy.special import logsumexp

import cases
from oracle import variational as OV
from oracle import distributions as OD

pytestmark = pytest.mark.gpu


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def zs():
    import zhusuan_b200 as zs
    return zs


def test_log_mean_exp_golden(zs):
    a = cases.LME_A
    for keepdims in [True, False]:
        true = logsumexp(a, (0, 2), keepdims=keepdims) - np.log(
            a.shape[0] * a.shape[2])
        np.testing.assert_allclose(N(zs.log_mean_exp(T(a), (0, 2), keepdims)),
                                   true, rtol=1e-6)
        np.testing.assert_allclose(N(zs.log_sum_exp(T(a), (0, 2), keepdims)),
                                   logsumexp(a, (0, 2), keepdims=keepdims),
                                   rtol=1e-6)
    b = cases.LME_B
    assert np.abs(N(zs.log_mean_exp(T(b), 0, False)) - b).max() < 1e-6


@pytest.mark.parametrize("shape,axis", [((64, 4096), 0), ((5, 7, 3), 1),
                                        ((9, 33), -1), ((6, 4, 5), None),
                                        ((1, 10), 0), ((3, 0), 0)])
def test_lme_vs_oracle_fwd_bwd(zs, shape, axis):
    rng = np.random.RandomState(0)
    x = (5 * rng.standard_normal(shape)).astype(np.float32)
    xt = T(x).requires_grad_(True)
    y = zs.log_mean_exp(xt, axis)
    ref = OV.log_mean_exp(x, axis, dtype=np.float64)
    np.testing.assert_allclose(N(y), ref, rtol=1e-5, atol=1e-5)
    if x.size == 0:
        return
    y.sum().backward()
    if axis is None:
        w = OV.iw_grad_logw(x.reshape(-1), 0, np.float64).reshape(shape)
    else:
        w = OV.iw_grad_logw(x, axis, np.float64)
    np.testing.assert_allclose(N(xt.grad), w, rtol=1e-4, atol=1e-6)


def _kl(m1, s1, m2, s2):
    return np.log(s2 / s1) + (s1 ** 2 + (m1 - m2) ** 2) / (2 * s2 ** 2) - 0.5


@pytest.mark.parametrize("x_mean,x_std", [(0., 1.), (2., 3.)])
def test_elbo_value_reference_test(zs, x_mean, x_std):
    """tests/variational/test_exclusive_kl.py:26-47 through the drop-in API."""
    z = np.random.RandomState(1).standard_normal(100000).astype(np.float32)
    log_q = stats.norm.logpdf(z).astype(np.float32)

    def log_joint(observed):
        return zs.distributions.Normal(mean=x_mean, std=x_std).log_prob(
            observed['x'])
    with pytest.warns(FutureWarning):
        lb = zs.variational.elbo(log_joint, observed={},
                                 latent={'x': [T(z), T(log_q)]}, axis=0)
    assert abs(float(lb.tensor) - (-_kl(0., 1., x_mean, x_std))) < 1e-3
    ref = OV.elbo(OD.normal_log_prob(z, x_mean, np.log(x_std)), [log_q], 0)
    np.testing.assert_allclose(float(lb.tensor), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("x_mean,x_std,rtol,atol",
                         [(0., 1., 1e-2, 1e-6), (2., 3., 1e-6, 1e-2)])
def test_elbo_sgvb_reference_test(zs, x_mean, x_std, rtol, atol):
    """tests/variational/test_exclusive_kl.py:49-78."""
    eps = T(np.random.RandomState(1).standard_normal(100000))
    mu = T(2.).requires_grad_(True)
    sigma = T(3.).requires_grad_(True)
    qx = eps * sigma + mu
    log_qx = zs.distributions.Normal(mean=mu, std=sigma).log_prob(qx)

    def log_joint(observed):
        return zs.distributions.Normal(mean=x_mean, std=x_std).log_prob(
            observed['x'])
    with pytest.warns(FutureWarning):
        lb = zs.variational.elbo(log_joint, observed={},
                                 latent={'x': [qx, log_qx]}, axis=0)
    g = torch.autograd.grad(lb.sgvb(), [mu, sigma])
    true = ((2. - x_mean) / x_std ** 2, -1 / 3. + 3. / x_std ** 2)
    np.testing.assert_allclose([float(g[0]), float(g[1])], true, rtol=rtol,
                               atol=atol)


@pytest.mark.parametrize("x_mean,x_std,thr", [(0., 1., 0.04), (2., 3., 0.02)])
def test_iwae_reference_tests(zs, x_mean, x_std, thr):
    """tests/variational/test_monte_carlo.py:25-102."""
    rng = np.random.RandomState(1)
    n1 = rng.standard_normal(size=(1, 1000)).astype(np.float32)
    n3 = rng.standard_normal(1000).astype(np.float32)

    def log_joint(observed):
        return zs.distributions.Normal(mean=x_mean, std=x_std).log_prob(
            observed['x'])
    analytic = -_kl(0., 1., x_mean, x_std)
    with pytest.warns(FutureWarning):
        lb = zs.variational.importance_weighted_objective(
            log_joint, observed={},
            latent={'x': [T(n1), T(stats.norm.logpdf(n1))]}, axis=0)
        lb3 = zs.variational.iw_objective(
            log_joint, observed={},
            latent={'x': [T(n3), T(stats.norm.logpdf(n3))]}, axis=0)
    assert abs(float(torch.mean(lb)) - analytic) < 1e-2
    assert float(torch.mean(lb3.tensor)) > analytic - 1e-6
    with pytest.raises(ValueError, match="axis"):
        zs.variational.iw_objective(log_joint, observed={},
                                    latent={'x': [T(n1), T(n1)]})
    # sgvb gradient
    mu = T(2.).requires_grad_(True)
    sigma = T(3.).requires_grad_(True)
    qx = T(n1) * sigma + mu
    log_qx = zs.distributions.Normal(mean=mu, std=sigma).log_prob(qx)
    with pytest.warns(FutureWarning):
        lb = zs.variational.iw_objective(log_joint, observed={},
                                         latent={'x': [qx, log_qx]}, axis=0)
    g = torch.autograd.grad(lb.sgvb().mean(), [mu, sigma])
    true = ((2. - x_mean) / x_std ** 2, -1 / 3. + 3. / x_std ** 2)
    np.testing.assert_allclose([float(g[0]), float(g[1])], true, rtol=thr,
                               atol=thr)


def test_iwae_vae_shaped_bayesian_net(zs):
    """config-3 wiring at a small size: q_net / gen BayesianNets, K particles,
    iw_objective(axis=0) value and SGVB gradient vs oracle + torch reference
    (examples/variational_autoencoders/iwae.py:23-75)."""
    torch.manual_seed(0)
    K, Nb, xd, zd, hd = 8, 16, 24, 5, 12
    dev = "cuda"
    W = {k: (0.3 * torch.randn(*s, device=dev)).requires_grad_(True)
         for k, s in dict(e1=(xd, hd), em=(hd, zd), es=(hd, zd),
                          d1=(zd, hd), d2=(hd, xd)).items()}
    x = (torch.rand(Nb, xd, device=dev) < 0.3).to(torch.int32)
    eps = torch.randn(K, Nb, zd, device=dev)

    @zs.meta_bayesian_net(scope="gen", reuse_variables=True)
    def build_gen(n, n_particles):
        bn = zs.BayesianNet()
        z = bn.normal("z", torch.zeros(n, zd, device=dev), std=1.,
                      group_ndims=1, n_samples=n_particles)
        h = torch.relu(z.tensor @ W["d1"])
        bn.bernoulli("x", h @ W["d2"], group_ndims=1)
        return bn

    def build_q_net(x, n_particles):
        bn = zs.BayesianNet()
        h = torch.relu(x.float() @ W["e1"])
        dist = zs.distributions.Normal(h @ W["em"], logstd=h @ W["es"],
                                       group_ndims=1)
        # inject eps so the oracle can replay the same draw
        node = bn.stochastic("z", dist, n_samples=n_particles)
        node._samples = dist._sample(n_particles, eps=eps)
        return bn

    model = build_gen(Nb, K)
    variational = build_q_net(x, K)
    lb = zs.variational.iw_objective(model, {'x': x}, variational=variational,
                                     axis=0)
    assert tuple(lb.tensor.shape) == (Nb,)
    cost = torch.mean(lb.sgvb())
    grads = torch.autograd.grad(cost, list(W.values()))

    # torch float64 reference of the same graph
    Wd = {k: v.detach().double().requires_grad_(True) for k, v in W.items()}
    xf = x.double()
    h = torch.relu(xf @ Wd["e1"])
    zm, zl = h @ Wd["em"], h @ Wd["es"]
    z = zm + torch.exp(zl) * eps.double()
    c = -0.5 * np.log(2 * np.pi)
    log_q = (c - zl - 0.5 * torch.exp(-2 * zl) * (z - zm) ** 2).sum(-1)
    log_pz = (c - 0.5 * z ** 2).sum(-1)
    logits = torch.relu(z @ Wd["d1"]) @ Wd["d2"]
    log_px = -(torch.clamp(logits, min=0) - logits * xf
               + torch.log1p(torch.exp(-logits.abs()))).sum(-1)
    log_w = log_pz + log_px - log_q
    ref = torch.logsumexp(log_w, 0) - np.log(K)
    np.testing.assert_allclose(N(lb.tensor), ref.detach().cpu().numpy(),
                               rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(
        N(lb.tensor), OV.iw_objective(log_w.detach().cpu().numpy(), [], 0,
                                      np.float64), rtol=1e-5, atol=1e-4)
    rg = torch.autograd.grad((-ref).mean(), list(Wd.values()))
    for a, b in zip(grads, rg):
        np.testing.assert_allclose(N(a), b.cpu().numpy(), rtol=2e-3, atol=2e-4)
    # is_loglikelihood alias (evaluation.py:22-54)
    ll = zs.is_loglikelihood(model, {'x': x}, axis=0, proposal=variational)
    np.testing.assert_allclose(N(ll), N(lb.tensor), rtol=1e-6)
