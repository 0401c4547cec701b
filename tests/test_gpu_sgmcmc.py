"""GPU parity: SG-MCMC samplers through the drop-in API vs the oracle-generated
golden vectors (tests/golden/sgmcmc.npz, injected noise)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def zs():
    import zhusuan_b200 as zs
    return zs


CONFIGS = {
    "sgld": ("SGLD", dict(learning_rate=0.01)),
    "psgld": ("PSGLD", dict(learning_rate=0.01)),
    "sghmc1": ("SGHMC", dict(learning_rate=0.01, friction=0.3,
                             variance_estimate=0.02, n_iter_resample_v=3,
                             second_order=False)),
    "sghmc2": ("SGHMC", dict(learning_rate=0.01, friction=0.3,
                             variance_estimate=0.02, n_iter_resample_v=3,
                             second_order=True)),
    "sgnht1v": ("SGNHT", dict(learning_rate=0.01, variance_extra=0.1,
                              tune_rate=2., n_iter_resample_v=4,
                              second_order=False, use_vector_alpha=True)),
    "sgnht2v": ("SGNHT", dict(learning_rate=0.01, variance_extra=0.1,
                              tune_rate=2., n_iter_resample_v=4,
                              second_order=True, use_vector_alpha=True)),
    "sgnht1s": ("SGNHT", dict(learning_rate=0.01, variance_extra=0.1,
                              tune_rate=2., n_iter_resample_v=None,
                              second_order=False, use_vector_alpha=False)),
    "sgnht2s": ("SGNHT", dict(learning_rate=0.01, variance_extra=0.1,
                              tune_rate=2., n_iter_resample_v=None,
                              second_order=True, use_vector_alpha=False)),
}


# sgmcmc.npz: written by the oracle; ref_sgmcmc.npz: written by the reference's own sgmcmc.py
# (oracle/tf_shim/make_ref_golden.py) -- same configurations, same keys
@pytest.mark.parametrize("fixture", ["sgmcmc.npz", "ref_sgmcmc.npz"])
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_golden_replay(zs, name, fixture):
    g = np.load(os.path.join(GOLD, fixture))
    cls, kw = CONFIGS[name]
    mean, std = T(g["mean"]), T(g["std"])

    @zs.meta_bayesian_net()
    def model():
        bn = zs.BayesianNet()
        bn.normal('x', mean, std=std, group_ndims=1)
        return bn
    x = T(g["q0"])
    s = getattr(zs, cls)(**kw)
    op, info = s.sample(model(), {}, {"x": x})
    if hasattr(s, "init_momentum"):
        s.init_momentum({"x": T(g[name + "_v0"])})
    for t in range(g[name + "_q"].shape[0]):
        op(noise={"noise": {"x": T(g[name + "_noise"][t])},
                  "resample": {"x": T(g[name + "_resample"][t])}})
        np.testing.assert_allclose(N(x), g[name + "_q"][t], rtol=1e-5,
                                   atol=1e-6)
        if name + "_mean_k" in g.files:
            np.testing.assert_allclose(N(info.mean_k["x"]),
                                       g[name + "_mean_k"][t], rtol=1e-4,
                                       atol=1e-7)
        if name + "_alpha" in g.files:
            np.testing.assert_allclose(N(info.alpha["x"]).reshape(-1),
                                       g[name + "_alpha"][t].reshape(-1),
                                       rtol=1e-4, atol=1e-6)
    assert info.q["x"] is x
    assert s.t == g[name + "_q"].shape[0]


def test_sghmc_double_well_statistical(zs):
    """tests/test_mcmc.py:72-79 on the GPU path with in-kernel Philox noise:
    SGHMC 1st order, 100 chains x 8000 iters, KDE error <= 0.016 (x1.5 slack:
    the reference's bound is '6 sd from 10 runs' of ITS rng stream)."""
    from scipy import stats
    x = torch.zeros(100, device="cuda")
    s = zs.SGHMC(learning_rate=0.01, n_iter_resample_v=50, friction=0.3,
                 variance_estimate=0.02, second_order=False, seed=5)

    def log_joint(obs):
        v = obs['x']
        return 2 * v ** 2 - v ** 4
    op, _ = s.sample(log_joint, {}, {"x": x})
    samples = []
    n_iters = 8000
    for t in range(n_iters):
        op()
        if t >= n_iters * 2 // 3 and t % 50 == 0:
            samples.append(N(x).copy())
    samples = np.array(samples).reshape(-1)
    assert not np.isnan(samples.sum())
    A = 3
    xs = np.linspace(-A, A, 1000)
    pdfs = np.exp(2 * (xs ** 2) - xs ** 4)
    pdfs = pdfs / pdfs.mean() / A / 2
    err = np.abs(stats.gaussian_kde(samples)(xs) - pdfs).mean()
    assert err <= 0.024


def test_type_error_contract(zs):
    s = zs.SGLD(learning_rate=0.1)
    with pytest.raises(TypeError, match=r"latent\['x'\] is not a"):
        s.sample(lambda o: o['x'].sum(-1), {}, {"x": [1.0, 2.0]})


@pytest.mark.parametrize("second_order", [True, False])
def test_bnn_fused_step_vs_oracle(zs, second_order):
    """config-4 shape at a small size: fused SGHMC+BNN kernel vs the float64
    oracle (oracle/models.py::BNN analytic gradient + oracle/sgmcmc.py::SGHMC)
    with injected noise, including the resample steps (t = 0, 3)."""
    from oracle import models as OM, sgmcmc as OS
    rng = np.random.RandomState(7)
    C, n_in, H, B, n_train = 9, 4, 37, 23, 500
    x = rng.standard_normal((B, n_in)); y = rng.standard_normal(B)
    ls0 = 0.1 * rng.standard_normal((H, n_in + 1)); ls1 = 0.1 * rng.standard_normal((1, H + 1))
    w0 = rng.uniform(-2, 2, (C, H, n_in + 1)); w1 = rng.uniform(-2, 2, (C, 1, H + 1))

    class M(OM.BNN):
        def grad(self, qs):
            g0, g1 = OM.BNN.grad(self, qs)
            # per-weight prior log-stddevs (bnn_sgmcmc.py:71)
            g0 = g0 + np.exp(-2 * self.ls0) * qs[0] - np.exp(-2 * ls0) * qs[0]
            g1 = g1 + np.exp(-2 * self.ls1) * qs[1] - np.exp(-2 * ls1) * qs[1]
            return [g0, g1]
    om = M(x, y, n_train, dtype=np.float64)
    kw = dict(learning_rate=1e-4, friction=0.2, variance_estimate=0.01,
              n_iter_resample_v=3, second_order=second_order)
    osg = OS.SGHMC(dtype=np.float64, **kw)
    v0n = [rng.standard_normal(w0.shape), rng.standard_normal(w1.shape)]
    osg.init_v(v0n)
    lj = zs.fused.BNNRegressionLogJoint(T(x), T(y), [T(ls0), T(ls1)], n_train)
    tw0, tw1 = T(w0), T(w1)
    sg = zs.SGHMC(**kw)
    op, info = sg.sample(lj, {}, {"w0": tw0, "w1": tw1})
    assert sg._fused_bnn() is lj
    sg.init_momentum({"w0": T(v0n[0]), "w1": T(v0n[1])})
    oq = [w0, w1]
    for t in range(5):
        nz = [rng.standard_normal(w0.shape), rng.standard_normal(w1.shape)]
        rs = [rng.standard_normal(w0.shape), rng.standard_normal(w1.shape)]
        oq, oinfo = osg.step(oq, om.grad, rs, nz)
        op(noise={"noise": {"w0": T(nz[0]), "w1": T(nz[1])},
                  "resample": {"w0": T(rs[0]), "w1": T(rs[1])}})
        np.testing.assert_allclose(N(tw0), oq[0], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(N(tw1), oq[1], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(N(info.mean_k["w0"]), oinfo["mean_k"][0], rtol=1e-3)
        np.testing.assert_allclose(N(info.mean_k["w1"]), oinfo["mean_k"][1], rtol=1e-3)


def test_bnn_fused_matches_generic_path_with_philox(zs):
    """[10, 50, 1] (bnn_sgmcmc.py) with in-kernel noise: the fused kernel and
    the generic path (registry kernels + autograd) draw the same Philox numbers
    and must agree step for step."""
    torch.manual_seed(0)
    C, n_in, H, B = 64, 10, 50, 100
    x = torch.randn(B, n_in, device="cuda"); y = torch.sin(x.sum(1))
    ls = [torch.zeros(H, n_in + 1, device="cuda"), torch.zeros(1, H + 1, device="cuda")]
    w0i = torch.rand(C, H, n_in + 1, device="cuda") * 4 - 2
    w1i = torch.rand(C, 1, H + 1, device="cuda") * 4 - 2
    outs = []
    for fused in (True, False):
        lj = zs.fused.BNNRegressionLogJoint(x, y, ls, n_train=10000)
        w0, w1 = w0i.clone(), w1i.clone()
        sg = zs.SGHMC(learning_rate=2e-6, friction=0.2, n_iter_resample_v=1000,
                      second_order=True, seed=11, use_fused=fused)
        op, info = sg.sample(lj, {}, {"w0": w0, "w1": w1})
        # identical initial momentum
        sg.vs[0].copy_(torch.full_like(w0, 1e-3)); sg.vs[1].copy_(torch.full_like(w1, -1e-3))
        for t in range(4):
            op()
        outs.append((N(w0), N(w1), float(info.mean_k["w0"])))
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-3)


@pytest.mark.parametrize("name", ["sgld", "psgld", "sghmc", "sgnht_vector", "sgnht_scalar"])
def test_sgmcmc_state_dict_round_trip_resumes_bitwise(zs, name):
    """SGMCMC.state_dict() / load_state_dict(): a sampler rebuilt from (latents, state dict)
    continues exactly like the uninterrupted one (iteration counter = Philox stream position,
    momenta, RMSprop accumulator, thermostats), across a momentum-resampling boundary."""
    rng = np.random.RandomState(2)
    D, C = 10, 12
    std = torch.tensor((0.5 + rng.random_sample(D)).astype(np.float32), device="cuda")
    lj = lambda o: zs.distributions.Normal(torch.zeros(D, device="cuda"), std=std,
                                           group_ndims=1).log_prob(o['w'])
    make = {"sgld": lambda: zs.SGLD(0.01, seed=3),
            "psgld": lambda: zs.PSGLD(0.01, seed=3),
            "sghmc": lambda: zs.SGHMC(0.01, friction=0.3, n_iter_resample_v=4, seed=3),
            "sgnht_vector": lambda: zs.SGNHT(0.01, variance_extra=0.1, n_iter_resample_v=4,
                                             seed=3),
            "sgnht_scalar": lambda: zs.SGNHT(0.01, variance_extra=0.1, use_vector_alpha=False,
                                             seed=3)}[name]
    q0 = torch.tensor(rng.standard_normal((C, D)).astype(np.float32), device="cuda")
    wa = q0.clone()
    sa = make()
    opa, _ = sa.sample(lj, {}, {"w": wa})
    for _ in range(3):
        opa()
    ckpt, w_ckpt = sa.state_dict(), wa.clone()
    for _ in range(5):
        opa()
    wb = w_ckpt.clone()
    sb = make()
    opb, _ = sb.sample(lj, {}, {"w": wb})
    sb.load_state_dict(ckpt)
    for _ in range(5):
        opb()
    torch.cuda.synchronize()
    assert sa.t == sb.t == 8
    np.testing.assert_array_equal(wa.cpu().numpy(), wb.cpu().numpy())


@pytest.mark.parametrize("fused", [True, False])
def test_bnn_sghmc_matches_reference_run(zs, fused):
    """tests/golden/ref_bnn_sghmc.npz: config 4's model (bnn_sgmcmc.py:19-35, 74-77) run on the
    reference's own BayesianNet + SGHMC classes (oracle/tf_shim/make_ref_golden.py).  The fused
    one-launch kernel (csrc/sgmcmc_bnn.cu) and the generic path must follow it step for step."""
    g = np.load(os.path.join(GOLD, "ref_bnn_sghmc.npz"))
    lj = zs.fused.BNNRegressionLogJoint(T(g["x"]), T(g["y"]), [T(g["logstd0"]), T(g["logstd1"])],
                                        int(g["n_train"]))
    w0, w1 = T(g["w0_init"]), T(g["w1_init"])
    sg = zs.SGHMC(learning_rate=float(g["cfg_learning_rate"]), friction=float(g["cfg_friction"]),
                  variance_estimate=float(g["cfg_variance_estimate"]),
                  n_iter_resample_v=int(g["cfg_n_iter_resample_v"]), second_order=True,
                  use_fused=fused)
    op, info = sg.sample(lj, {}, {"w0": w0, "w1": w1})
    assert sg._fused_bnn() is lj          # the log-joint is recognised; use_fused picks the path
    sg.init_momentum({"w0": T(g["v0_0"]), "w1": T(g["v0_1"])})
    for t in range(g["w0"].shape[0]):
        op(noise={"noise": {"w0": T(g["noise0"][t]), "w1": T(g["noise1"][t])},
                  "resample": {"w0": T(g["resample0"][t]), "w1": T(g["resample1"][t])}})
        np.testing.assert_allclose(N(w0), g["w0"][t], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(N(w1), g["w1"][t], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(N(info.mean_k["w0"]), g["mean_k0"][t], rtol=1e-3)
        np.testing.assert_allclose(N(info.mean_k["w1"]), g["mean_k1"][t], rtol=1e-3)
