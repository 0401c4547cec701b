"""GPU tests of the remaining config shapes on the generic path: the LNTM
log-joint (config 5, examples/topic_models/lntm_mcem.py:33-48, 97-105) through
the BayesianNet contract incl. a two-chain-axis HMC run vs the oracle, and AIS
(evaluation.py:57-172) against an analytic marginal likelihood."""
import os

import numpy as np
import pytest
import torch

from oracle import hmc as OH
from oracle import models as OM

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


def test_ais_device_loop_matches_reference_run(zs):
    """tests/golden/ref_ais.npz: the reference's own class AIS (evaluation.py:57-172) and HMC,
    executed on the NumPy TF stand-in (oracle/tf_shim/make_ref_golden.py) with every draw
    injected; the device loop must reproduce its per-chain log-weights and bound."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ais.npz"))
    s = float(g["s"])
    n_chains, n_data, d = g["init"].shape[1:]
    nt, na = int(g["n_temperatures"]), int(g["n_adapt"])

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
    ais = zs.AIS(make(True), make(False), hmc, observed={'x': T(g["x"])}, latent={'z': z},
                 n_temperatures=nt, n_adapt=na)
    est = ais.run(noise=lambda k: {"p": {"z": T(g["noise_p"][k])}, "u": T(g["noise_u"][k])},
                  init=[[T(g["init"][0])], [T(g["init"][1])]])
    np.testing.assert_allclose(N(ais.log_weights), g["log_weights"], rtol=2e-4, atol=2e-4)
    assert abs(est - float(g["bound"])) < 2e-4
    np.testing.assert_allclose(N(z), g["z_final"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(N(ais._schedule), g["schedule"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("K,V,C,Dn", [(16, 50, 5, 3), (32, 200, 70, 4), (128, 1000, 130, 7)])
def test_lntm_fused_kernel_matches_oracle(zs, K, V, C, Dn):
    """zs.fused.LNTMLogJoint (sparsity-aware fused kernel, csrc/lntm.cu) vs the dense float64
    oracle restatement of lntm_mcem.py:33-48 / e_obj: log-joint values and d/d eta, including an
    empty (padding) document, counts > 1, and a chain count that is not a multiple of the block's
    64; the dense torch restatement (the object as a callable) agrees as well."""
    rng = np.random.RandomState(K + V)
    x = rng.poisson(0.08, (Dn, V)).astype(np.float32)
    x[0] = 0                                             # padding document (lntm_mcem.py:71-74)
    x[1, :5] += 3
    beta = rng.standard_normal((K, V)).astype(np.float32)
    mean = (0.3 * rng.standard_normal(K)).astype(np.float32)
    logstd = (0.2 * rng.standard_normal(K)).astype(np.float32)
    eta = rng.standard_normal((C, Dn, K)).astype(np.float32)
    om = OM.LNTM(x, beta, mean, logstd)
    lj = zs.fused.LNTMLogJoint(T(x), T(beta), T(mean), T(logstd))
    assert int(lj.doc_ptr[-1]) == int((x != 0).sum()) and int(lj.doc_ptr[1]) == 0
    lp = N(lj.logp([T(eta)]))
    g = N(lj.grad([T(eta)])[0])
    ref_lp, ref_g = om.logp([eta]), om.grad([eta])[0]
    np.testing.assert_allclose(lp, ref_lp, rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(g, ref_g, rtol=2e-4, atol=2e-4 * np.abs(ref_g).max())
    dense = N(lj({"eta": T(eta)}))
    np.testing.assert_allclose(dense, ref_lp, rtol=5e-5, atol=5e-3)


def test_lntm_fused_hmc_matches_oracle(zs):
    """HMC on the fused LNTM provider (two chain axes [chains, docs], lntm_mcem.py:69-70,
    99-105) vs the oracle HMC on the dense oracle model, injected noise, adaptive step size."""
    rng = np.random.RandomState(9)
    K, V, C, Dn = 32, 120, 6, 5
    x = rng.poisson(0.1, (Dn, V)).astype(np.float32)
    beta = rng.standard_normal((K, V)).astype(np.float32)
    mean = np.zeros(K, np.float32)
    logstd = np.zeros(K, np.float32)
    om = OM.LNTM(x, beta, mean, logstd, dtype=np.float32)
    lj = zs.fused.LNTMLogJoint(T(x), T(beta), T(mean), T(logstd))
    eta0 = (0.1 * rng.standard_normal((C, Dn, K))).astype(np.float32)
    eta = T(eta0)
    h = zs.HMC(step_size=0.02, n_leapfrogs=6, adapt_step_size=True, target_acceptance_rate=0.6)
    op, info = h.sample(lj, {}, {"eta": eta})
    assert h._provider is lj and tuple(info.acceptance_rate.shape) == (C, Dn)
    oh = OH.HMC(step_size=0.02, n_leapfrogs=6, adapt_step_size=True, target_acceptance_rate=0.6)
    oq = [eta0]
    for i in range(4):
        npz = rng.standard_normal(eta0.shape).astype(np.float32)
        u = rng.random_sample((C, Dn)).astype(np.float32)
        with np.errstate(all="ignore"):
            oq, oi = oh.step(oq, om.logp, om.grad, [npz], u, True, False)
        op(adapt_step_size=True, noise={"p": {"eta": T(npz)}, "u": T(u)})
        np.testing.assert_allclose(N(info.acceptance_rate), oi.acceptance_rate, rtol=2e-3,
                                   atol=2e-4)
        near = np.abs(u - oi.acceptance_rate) < 1e-3
        np.testing.assert_allclose(N(eta)[~near], oq[0][~near], rtol=1e-3, atol=1e-4)


def test_lntm_fused_hmc_follows_reference_run(zs):
    """tests/golden/ref_lntm_hmc.npz: config 5's E-step on the reference's own BayesianNet /
    UnnormalizedMultinomial / HMC (oracle/tf_shim/make_ref_golden.py).  HMC on the fused
    sparsity-aware LNTM kernel (csrc/lntm.cu) follows it iteration by iteration."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "ref_lntm_hmc.npz"))
    lj = zs.fused.LNTMLogJoint(T(g["x"]), T(g["beta"]), T(g["eta_mean"]), T(g["eta_logstd"]))
    eta = T(g["eta0"])
    h = zs.HMC(step_size=float(g["cfg_step_size"]), n_leapfrogs=int(g["cfg_n_leapfrogs"]),
               adapt_step_size=True, target_acceptance_rate=float(g["cfg_target_acceptance_rate"]))
    op, info = h.sample(lj, {}, {"eta": eta})
    assert h._provider is lj
    for i in range(g["eta"].shape[0]):
        op(adapt_step_size=True, noise={"p": {"eta": T(g["noise_p"][i])}, "u": T(g["noise_u"][i])})
        np.testing.assert_allclose(N(info.orig_log_prob), g["lp0"][i], rtol=2e-5, atol=2e-3)
        np.testing.assert_allclose(N(info.acceptance_rate), g["acc"][i], rtol=3e-3, atol=3e-4)
        np.testing.assert_allclose(float(info.updated_step_size), g["step_size"][i], rtol=3e-4)
        near = np.abs(g["noise_u"][i] - g["acc"][i]) < 2e-3
        np.testing.assert_allclose(N(eta)[~near], g["eta"][i][~near], rtol=1e-3, atol=2e-4)
        eta.copy_(T(g["eta"][i]))            # continue from the reference's state
