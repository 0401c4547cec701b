"""GPU tests of the remaining config shapes on the generic path: the LNTM
log-joint (config 5, examples/topic_models/lntm_mcem.py:33-48, 97-105) through
the BayesianNet contract incl. a two-chain-axis HMC run vs the oracle, and AIS
(evaluation.py:57-172) against an analytic marginal likelihood."""
import numpy as np
import pytest
import torch

from oracle import hmc as OH

pytestmark = pytest.mark.gpu


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def zs():
    import zhusuan_b200 as zs
    return zs


def _lntm_reference(eta, x, beta, eta_mean, eta_logstd):
    """float64 torch-CPU restatement of lntm_mcem.py:33-48 with e_obj (97-98)."""
    eta = eta.double()
    c = -0.5 * np.log(2 * np.pi)
    lp_eta = (c - eta_logstd - 0.5 * torch.exp(-2 * eta_logstd)
              * (eta - eta_mean) ** 2).sum(-1)
    theta = torch.softmax(eta, -1)
    phi = torch.softmax(beta, -1)
    doc_word = theta @ phi
    lp_x = (x * torch.log(doc_word)).sum(-1)
    return lp_eta + lp_x


def test_lntm_log_joint_gradient_and_hmc(zs):
    rng = np.random.RandomState(0)
    n_chains, n_docs, n_topics, n_vocab = 3, 5, 6, 40
    beta = rng.standard_normal((n_topics, n_vocab))
    x = rng.poisson(0.5, (n_docs, n_vocab)).astype(np.float32)
    eta_mean = 0.1 * rng.standard_normal(n_topics)
    eta_logstd = 0.2 * rng.standard_normal(n_topics)
    eta0 = 0.3 * rng.standard_normal((n_chains, n_docs, n_topics)).astype(np.float32)
    tb, tx, tm, tl = T(beta), T(x), T(eta_mean), T(eta_logstd)

    @zs.meta_bayesian_net(scope='lntm')
    def lntm(n_chains, n_docs, n_topics, n_vocab, eta_mean, eta_logstd):
        bn = zs.BayesianNet()
        em = eta_mean.unsqueeze(0).expand(n_docs, -1)
        eta = bn.normal('eta', em, logstd=eta_logstd, n_samples=n_chains,
                        group_ndims=1)
        theta = torch.softmax(eta.tensor, -1)
        beta = bn.normal('beta', torch.zeros(n_topics, n_vocab, device="cuda"),
                         logstd=10.0, group_ndims=1)
        phi = torch.softmax(beta.tensor, -1)
        doc_word = (theta.reshape(-1, n_topics) @ phi).reshape(
            n_chains, n_docs, n_vocab)
        bn.unnormalized_multinomial('x', torch.log(doc_word),
                                    normalize_logits=False,
                                    dtype=torch.float32)
        return bn
    model = lntm(n_chains, n_docs, n_topics, n_vocab, tm, tl)
    model.log_joint = lambda bn: bn.cond_log_prob('eta') + bn.cond_log_prob('x')

    # value + gradient of the log-joint vs float64
    eta = T(eta0).requires_grad_(True)
    lj = model.observe(eta=eta, x=tx, beta=tb).log_joint()
    assert tuple(lj.shape) == (n_chains, n_docs)
    e64 = torch.tensor(eta0, dtype=torch.float64, requires_grad=True)
    ref = _lntm_reference(e64, torch.tensor(x, dtype=torch.float64),
                          torch.tensor(beta), torch.tensor(eta_mean),
                          torch.tensor(eta_logstd))
    np.testing.assert_allclose(N(lj), ref.detach().numpy(), rtol=1e-5, atol=1e-4)
    lj.sum().backward(); ref.sum().backward()
    np.testing.assert_allclose(N(eta.grad), e64.grad.numpy(), rtol=1e-4, atol=1e-4)

    # HMC with TWO chain axes [n_chains, n_docs] (lntm_mcem.py:69-70, 99-105)
    def logp(q):
        with torch.no_grad():
            return _lntm_reference(torch.tensor(q[0]), torch.tensor(x, dtype=torch.float64),
                                   torch.tensor(beta), torch.tensor(eta_mean),
                                   torch.tensor(eta_logstd)).numpy().astype(np.float32)

    def grad(q):
        e = torch.tensor(q[0], dtype=torch.float64, requires_grad=True)
        _lntm_reference(e, torch.tensor(x, dtype=torch.float64), torch.tensor(beta),
                        torch.tensor(eta_mean), torch.tensor(eta_logstd)).sum().backward()
        return [e.grad.numpy().astype(np.float32)]
    oh = OH.HMC(step_size=0.05, n_leapfrogs=5, adapt_step_size=True,
                target_acceptance_rate=0.6)
    h = zs.HMC(step_size=0.05, n_leapfrogs=5, adapt_step_size=True,
               target_acceptance_rate=0.6)
    eta_v = T(eta0)
    op, info = h.sample(model, observed={'x': tx, 'beta': tb}, latent={'eta': eta_v})
    assert tuple(info.acceptance_rate.shape) == (n_chains, n_docs)
    oq = [eta0]
    for i in range(4):
        npz = rng.standard_normal(eta0.shape).astype(np.float32)
        u = rng.random_sample((n_chains, n_docs)).astype(np.float32)
        oq, oi = oh.step(oq, logp, grad, [npz], u, True, False)
        op(adapt_step_size=True, noise={"p": {"eta": T(npz)}, "u": T(u)})
        np.testing.assert_allclose(N(info.acceptance_rate), oi.acceptance_rate,
                                   rtol=2e-3, atol=2e-4)
        np.testing.assert_allclose(N(eta_v), oq[0], rtol=1e-3, atol=1e-4)


def test_ais_recovers_analytic_marginal_likelihood(zs):
    """z ~ N(0, I_d), x | z ~ N(z, s^2 I_d)  =>  log p(x) = log N(x; 0, (1+s^2) I).
    AIS (HMC transitions, sigmoid schedule) must land within 0.1 nat per datum."""
    from scipy import stats
    torch.manual_seed(0)
    zs.set_random_seed(3)
    n_chains, n_data, d, s = 32, 6, 3, 0.7
    x = T(np.random.RandomState(1).standard_normal((n_data, d)) * np.sqrt(1 + s * s))

    def make(include_x):
        @zs.meta_bayesian_net()
        def m():
            bn = zs.BayesianNet()
            z = bn.normal('z', torch.zeros(n_data, d, device="cuda"), std=1.,
                          group_ndims=1, n_samples=n_chains)
            if include_x:
                bn.normal('x', z.tensor, std=s, group_ndims=1)
            return bn
        return m()
    model, proposal = make(True), make(False)
    hmc = zs.HMC(step_size=0.05, n_leapfrogs=8, adapt_step_size=True,
                 target_acceptance_rate=0.7, seed=5)
    z = torch.zeros(n_chains, n_data, d, device="cuda")
    ais = zs.AIS(model, proposal, hmc, observed={'x': x}, latent={'z': z},
                 n_temperatures=120, n_adapt=20)
    est = ais.run()
    truth = stats.norm.logpdf(N(x), 0, np.sqrt(1 + s * s)).sum(-1).mean()
    assert abs(est - truth) < 0.1, (est, truth)
    assert tuple(ais.log_weights.shape) == (n_chains, n_data)


def test_ais_device_loop_matches_oracle_step_for_step(zs):
    """evaluation.py:119-172 with every draw injected (prior samples, HMC noise): the device
    loop's per-chain log-weights and bound equal the oracle restatement (oracle/evaluation.py);
    a feed_dict given to run() reaches the HMC transitions, not only the prior density."""
    from oracle import evaluation as OE
    rng = np.random.RandomState(4)
    n_chains, n_data, d, s = 8, 3, 2, 0.8
    nt, na = 12, 3
    x_np = (rng.standard_normal((n_data, d)) * 1.2).astype(np.float32)
    x_wrong = torch.zeros(n_data, d, device="cuda")     # construction-time value, replaced
    obs = {'x': x_wrong}

    def make(include_x):
        @zs.meta_bayesian_net()
        def m():
            bn = zs.BayesianNet()
            z = bn.normal('z', torch.zeros(n_data, d, device="cuda"), std=1.,
                          group_ndims=1, n_samples=n_chains)
            if include_x:
                bn.normal('x', z.tensor, std=s, group_ndims=1)
            return bn
        return m()
    hmc = zs.HMC(step_size=0.2, n_leapfrogs=3, adapt_step_size=True,
                 target_acceptance_rate=0.7)
    z = torch.zeros(n_chains, n_data, d, device="cuda")
    ais = zs.AIS(make(True), make(False), hmc, observed=obs, latent={'z': z},
                 n_temperatures=nt, n_adapt=na)
    init = [rng.standard_normal((n_chains, n_data, d)).astype(np.float32) for _ in range(2)]
    noises = [(rng.standard_normal((n_chains, n_data, d)).astype(np.float32),
               rng.random_sample((n_chains, n_data)).astype(np.float32))
              for _ in range(na + nt)]
    est = ais.run(feed_dict={'x': T(x_np)},
                  noise=lambda k: {"p": {"z": T(noises[k][0])}, "u": T(noises[k][1])},
                  init=[[T(init[0])], [T(init[1])]])

    c = -0.5 * np.log(2 * np.pi)
    f32 = np.float32
    lp = lambda q: (c - 0.5 * q[0].astype(np.float64) ** 2).sum(-1).astype(f32)
    gp = lambda q: [(-q[0]).astype(f32)]
    lj = lambda q: (lp(q).astype(np.float64) + (c - np.log(s) - 0.5 * (
        (x_np - q[0].astype(np.float64)) / s) ** 2).sum(-1)).astype(f32)
    gj = lambda q: [(-q[0] + (x_np - q[0]) / (s * s)).astype(f32)]
    oh = OH.HMC(step_size=0.2, n_leapfrogs=3, adapt_step_size=True, target_acceptance_rate=0.7)
    oa = OE.AIS(lp, gp, lj, gj, oh, n_temperatures=nt, n_adapt=na)
    oest, olw = oa.run([[init[0]], [init[1]]], lambda k: ([noises[k][0]], noises[k][1]),
                       adapt_flags=(True, False))
    np.testing.assert_allclose(N(ais.log_weights), olw, rtol=2e-4, atol=2e-4)
    assert abs(est - oest) < 2e-4
    np.testing.assert_allclose(N(ais._schedule), [oa.schedule(t) for t in range(nt + 1)],
                               rtol=1e-6, atol=1e-7)
    assert ais.temperature.is_cuda and float(ais.temperature) == 1.0
