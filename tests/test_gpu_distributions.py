"""GPU parity: distribution log_prob / backward / sampling kernels vs the
reference's known answers (tests/cases.py) and vs the CPU oracle.
Tolerance: float32 path, rtol 1e-5 (north_star: "within 1e-5 relative fp32")."""
import numpy as np
import pytest
import torch

import cases
from oracle import distributions as OD
from oracle import philox

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-5


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def zs():
    import zhusuan_b200 as zs
    return zs


def test_normal_known_answers(zs):
    for given, mean, logstd, g, target in cases.normal_cases():
        d = zs.distributions.Normal(T(mean), logstd=T(logstd), group_ndims=g)
        np.testing.assert_allclose(N(d.log_prob(T(given))), target,
                                   rtol=RTOL, atol=ATOL)
        d2 = zs.distributions.Normal(T(mean), std=T(np.exp(logstd)),
                                     group_ndims=g)
        np.testing.assert_allclose(N(d2.log_prob(T(given))), target,
                                   rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(N(d.prob(T(given))), np.exp(target),
                                   rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize("shape,pshape,g", [
    ((64, 4096 // 64, 40), (4096 // 64, 40), 1),     # particles x batch x z
    ((7, 5, 3), (1, 5, 3), 2), ((33, 100), (100,), 1), ((1000,), (), 0),
    ((4, 6, 5), (6, 1), 1), ((0, 5), (5,), 1)])
def test_normal_vs_oracle_and_grads(zs, shape, pshape, g):
    rng = np.random.RandomState(0)
    x = rng.standard_normal(shape).astype(np.float32)
    mu = rng.standard_normal(pshape).astype(np.float32)
    ls = (0.3 * rng.standard_normal(pshape)).astype(np.float32)
    xt, mt, lt = (T(a).requires_grad_(True) for a in (x, mu, ls))
    lp = zs.distributions.Normal(mt, logstd=lt, group_ndims=g).log_prob(xt)
    ref = OD.normal_log_prob(x, mu, ls, g, np.float64)
    assert tuple(lp.shape) == ref.shape
    np.testing.assert_allclose(N(lp), ref, rtol=RTOL, atol=1e-4)
    if x.size == 0:
        return
    w = rng.standard_normal(ref.shape).astype(np.float32)
    (lp * T(w)).sum().backward()
    dg, dm, dl = OD.normal_log_prob_grads(x, mu, ls)
    wexp = np.broadcast_to(w.reshape(w.shape + (1,) * g), dg.shape) \
        if g else np.broadcast_to(w, dg.shape)

    def red(a, s):
        a = a * wexp
        while a.ndim > len(s):
            a = a.sum(0)
        for ax, n in enumerate(s):
            if n == 1 and a.shape[ax] != 1:
                a = a.sum(ax, keepdims=True)
        return a
    np.testing.assert_allclose(N(xt.grad), red(dg, x.shape), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(N(mt.grad), red(dm, mu.shape), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(N(lt.grad), red(dl, ls.shape), rtol=1e-4, atol=1e-3)


def test_bernoulli_known_answers_and_grad(zs):
    for logits, given, target in cases.bernoulli_cases():
        d = zs.distributions.Bernoulli(T(logits))
        np.testing.assert_allclose(N(d.log_prob(T(given, torch.int32))),
                                   target, rtol=RTOL, atol=ATOL)
    # config-3 shaped: logits [K, N, 784] vs x [N, 784], group_ndims=1
    rng = np.random.RandomState(1)
    l = (3 * rng.standard_normal((4, 6, 784))).astype(np.float32)
    x = (rng.random_sample((6, 784)) < 0.13).astype(np.int32)
    lt = T(l).requires_grad_(True)
    lp = zs.distributions.Bernoulli(lt, group_ndims=1).log_prob(
        T(x, torch.int32))
    np.testing.assert_allclose(N(lp), OD.bernoulli_log_prob(x, l, 1, np.float64),
                               rtol=RTOL, atol=1e-3)
    lp.sum().backward()
    ref = x[None].astype(np.float64) - 1 / (1 + np.exp(-l.astype(np.float64)))
    np.testing.assert_allclose(N(lt.grad), ref, rtol=1e-4, atol=1e-5)


def test_categorical_known_answers_and_grad(zs):
    for logits, given, target in cases.categorical_cases():
        d = zs.distributions.Categorical(T(logits))
        np.testing.assert_allclose(N(d.log_prob(T(given, torch.int32))),
                                   target, rtol=RTOL, atol=ATOL)
    rng = np.random.RandomState(2)
    l = rng.standard_normal((5, 7, 11)).astype(np.float32)
    k = rng.randint(0, 11, size=(5, 7)).astype(np.int32)
    lt = T(l).requires_grad_(True)
    lp = zs.distributions.Categorical(lt, group_ndims=1).log_prob(
        T(k, torch.int32))
    np.testing.assert_allclose(N(lp), OD.categorical_log_prob(k, l, 1, np.float64),
                               rtol=RTOL, atol=1e-5)
    lp.sum().backward()
    sm = np.exp(l - l.max(-1, keepdims=True)); sm /= sm.sum(-1, keepdims=True)
    ref = np.eye(11)[k] - sm
    np.testing.assert_allclose(N(lt.grad), ref, rtol=1e-4, atol=1e-5)


def test_unnormalized_multinomial_known_answers_and_grad(zs):
    for logits, given, norm, target in cases.unnorm_multinomial_cases():
        d = zs.distributions.UnnormalizedMultinomial(T(logits),
                                                     normalize_logits=norm)
        np.testing.assert_allclose(N(d.log_prob(T(given, torch.int32))),
                                   target, rtol=1e-5, atol=1e-3)
    rng = np.random.RandomState(3)
    l = rng.standard_normal((3, 4, 50)).astype(np.float32)     # [chains, docs, V]
    x = rng.poisson(2.0, size=(4, 50)).astype(np.int32)        # [docs, V]
    lt = T(l).requires_grad_(True)
    lp = zs.distributions.UnnormalizedMultinomial(lt).log_prob(
        T(x, torch.int32))
    np.testing.assert_allclose(
        N(lp), OD.unnormalized_multinomial_log_prob(x, l, True, 0, np.float64),
        rtol=RTOL, atol=1e-3)
    lp.sum().backward()
    sm = np.exp(l - l.max(-1, keepdims=True)); sm /= sm.sum(-1, keepdims=True)
    ref = x[None] - x.sum(-1, keepdims=True)[None] * sm
    np.testing.assert_allclose(N(lt.grad), ref, rtol=1e-4, atol=1e-4)


def test_dirichlet_known_answers_and_grad(zs):
    for alpha, given, target in cases.dirichlet_cases():
        d = zs.distributions.Dirichlet(T(alpha))
        np.testing.assert_allclose(N(d.log_prob(T(given))), target,
                                   rtol=1e-5, atol=1e-5)
    a = np.array([[2., 3., 4.], [1.5, 0.7, 5.]], np.float32)
    x = np.array([[0.2, 0.3, 0.5], [0.1, 0.6, 0.3]], np.float32)
    xt = T(x).requires_grad_(True)
    zs.distributions.Dirichlet(T(a)).log_prob(xt).sum().backward()
    np.testing.assert_allclose(N(xt.grad), (a - 1) / x, rtol=1e-5)
    with pytest.raises(ValueError):
        zs.distributions.Dirichlet(T([1.0]))


@pytest.mark.parametrize("seed", [23, 233, 2333])
def test_mvn_cholesky_vs_scipy(zs, seed):
    from scipy import stats
    mean, cov, chol = cases.mvn_params(seed)
    rng = np.random.RandomState(seed)
    samples = mean + np.einsum('ijab,nijb->nija', chol,
                               rng.standard_normal((6,) + mean.shape))
    d = zs.distributions.MultivariateNormalCholesky(T(mean), T(chol))
    st = T(samples).requires_grad_(True)
    lp = d.log_prob(st)
    assert tuple(lp.shape) == (6,) + mean.shape[:2]
    for i in range(mean.shape[0]):
        for j in range(mean.shape[1]):
            exact = stats.multivariate_normal.logpdf(
                samples[:, i, j, :], mean[i, j], cov[i, j])
            np.testing.assert_allclose(N(lp)[:, i, j], exact, rtol=2e-4,
                                       atol=2e-3)
    lp.sum().backward()
    prec = np.linalg.inv(cov)
    ref = -np.einsum('ijab,nijb->nija', prec, samples - mean)
    np.testing.assert_allclose(N(st.grad), ref, rtol=2e-3, atol=2e-2)


def test_parameter_gradients_of_mvn_cholesky_and_dirichlet(zs):
    """ELBO / IWAE with a TRAINABLE MultivariateNormalCholesky or Dirichlet posterior
    differentiates log q w.r.t. cov_tril / alpha (the reference does it through TF autodiff of
    multivariate.py:169-189, 665-677): checked against float64 torch autograd of the same
    formulas, with cov_tril broadcast over a leading sample axis."""
    rng = np.random.RandomState(7)
    Dn = 5
    mean = rng.standard_normal((3, Dn))
    A = rng.standard_normal((3, Dn, Dn)) * 0.3
    chol = np.tril(A) + np.eye(Dn) * (1.0 + rng.random_sample((3, Dn, 1)) * 0.5) * np.eye(Dn)
    chol = np.tril(chol)
    x = rng.standard_normal((4, 3, Dn))
    w = rng.standard_normal((4, 3))
    # float64 reference
    m64 = torch.tensor(mean, dtype=torch.float64, requires_grad=True)
    c64 = torch.tensor(chol, dtype=torch.float64, requires_grad=True)
    x64 = torch.tensor(x, dtype=torch.float64)
    y = torch.linalg.solve_triangular(c64.expand(4, 3, Dn, Dn), (x64 - m64).unsqueeze(-1),
                                      upper=False).squeeze(-1)
    lp64 = (-0.5 * Dn * np.log(2 * np.pi) - torch.log(torch.diagonal(c64, dim1=-2, dim2=-1)).sum(-1)
            - 0.5 * (y * y).sum(-1))
    (lp64 * torch.tensor(w)).sum().backward()
    mt, ct = T(mean).requires_grad_(True), T(chol).requires_grad_(True)
    lp = zs.distributions.MultivariateNormalCholesky(mt, ct).log_prob(T(x))
    np.testing.assert_allclose(N(lp), lp64.detach().numpy(), rtol=1e-4, atol=1e-4)
    (lp * T(w)).sum().backward()
    np.testing.assert_allclose(N(ct.grad), c64.grad.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(N(mt.grad), m64.grad.numpy(), rtol=2e-4, atol=2e-4)
    assert np.all(np.triu(N(ct.grad), 1) == 0)

    a = 0.5 + 3 * rng.random_sample((3, 4))
    g = rng.dirichlet(np.ones(4), size=(5, 3))
    w2 = rng.standard_normal((5, 3))
    a64 = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    g64 = torch.tensor(g, dtype=torch.float64)
    lpd = (torch.lgamma(a64.sum(-1)) - torch.lgamma(a64).sum(-1)
           + ((a64 - 1) * torch.log(g64)).sum(-1))
    (lpd * torch.tensor(w2)).sum().backward()
    at = T(a).requires_grad_(True)
    lp = zs.distributions.Dirichlet(at).log_prob(T(g))
    np.testing.assert_allclose(N(lp), lpd.detach().numpy(), rtol=1e-4, atol=1e-4)
    (lp * T(w2)).sum().backward()
    np.testing.assert_allclose(N(at.grad), a64.grad.numpy(), rtol=2e-4, atol=2e-4)


def test_shape_and_dtype_contract(zs):
    d = zs.distributions.Normal(T(np.zeros((2, 3))), logstd=T(np.zeros(3)),
                                group_ndims=1)
    assert tuple(d.get_batch_shape()) == (2, 3)
    assert tuple(d.get_value_shape()) == ()
    assert tuple(d.sample(5).shape) == (5, 2, 3)
    assert tuple(d.sample().shape) == (2, 3)
    with pytest.raises(ValueError, match="broadcast to match"):
        d.log_prob(T(np.zeros((4, 5))))
    with pytest.raises(ValueError, match="Either `std` or `logstd`"):
        zs.distributions.Normal(0.)
    with pytest.raises(TypeError, match="must have the same dtype as"):
        zs.distributions.Normal(T(0.), std=T(1., torch.float64))
    with pytest.raises(ValueError, match="non-negative"):
        zs.distributions.Normal(0., std=1., group_ndims=-1)
    u = zs.distributions.UnnormalizedMultinomial(T(np.zeros((2, 3))))
    with pytest.raises(NotImplementedError):
        u.sample(2)
    b = zs.distributions.Bernoulli(T(np.zeros((2, 3))))
    s = b.sample(7)
    assert s.dtype == torch.int32 and tuple(s.shape) == (7, 2, 3)
    assert set(np.unique(N(s))) <= {0, 1}


def test_reparameterised_sample_injected_and_philox(zs):
    rng = np.random.RandomState(5)
    mu = rng.standard_normal((4, 40)).astype(np.float32)
    ls = (0.2 * rng.standard_normal((4, 40))).astype(np.float32)
    eps = rng.standard_normal((8, 4, 40)).astype(np.float32)
    mt, lt = T(mu).requires_grad_(True), T(ls).requires_grad_(True)
    d = zs.distributions.Normal(mt, logstd=lt, group_ndims=1)
    z = d._sample(8, eps=T(eps))
    np.testing.assert_allclose(N(z), OD.normal_sample(eps, mu, np.exp(ls)),
                               rtol=1e-6, atol=1e-6)
    (z * z).sum().backward()
    zz = OD.normal_sample(eps, mu, np.exp(ls), np.float64)
    np.testing.assert_allclose(N(mt.grad), (2 * zz).sum(0), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(N(lt.grad), (2 * zz * eps * np.exp(ls)).sum(0),
                               rtol=1e-4, atol=1e-3)
    # in-kernel Philox: standard-normal moments, reproducible per seed
    zs.set_random_seed(1234)
    d2 = zs.distributions.Normal(T(np.zeros(4096)), std=T(np.ones(4096)))
    s1 = N(d2.sample(64))
    assert abs(s1.mean()) < 0.01 and abs(s1.std() - 1) < 0.01
    zs.set_random_seed(1234)
    d3 = zs.distributions.Normal(T(np.zeros(4096)), std=T(np.ones(4096)))
    np.testing.assert_array_equal(s1, N(d3.sample(64)))


def test_kernel_philox_matches_oracle():
    """momentum kernel with mass = 1 exposes the in-kernel normals; the MH
    kernel exposes the uniforms (accept <=> u < acc with acc = 1/2...)."""
    import ctypes
    from zhusuan_b200._lib import lib, ptr, stream
    C, D = 37, 22
    p = torch.empty(C, D, device="cuda")
    mass = torch.ones(D, device="cuda")
    seed, it, row0 = 0x1234567887654321, 9, 1000
    lib.call("zsb_hmc_momentum_f32", ptr(p), None, ptr(mass), D, C, D, seed,
             it, 1, row0, None, 0, None, stream())
    ref = philox.normal_matrix(seed, 1, it, row0, C, D)
    np.testing.assert_allclose(N(p), ref, rtol=1e-5, atol=2e-6)
    # uniforms: lp1 - lp0 = log(t) -> acc = t; accept <=> u < t
    u_ref = philox.uniform_vector(seed, 2, it, row0, C)
    z = torch.zeros(C, device="cuda")
    thr = 0.5
    lp1 = torch.full((C,), float(np.log(thr)), device="cuda")
    acc = torch.empty(C, device="cuda")
    accept = torch.empty(C, dtype=torch.int32, device="cuda")
    part = torch.empty(lib.load().zsb_hmc_acc_parts(), device="cuda")
    npart = ctypes.c_int(0)
    lib.call("zsb_hmc_mh_f32", ptr(z), ptr(lp1), ptr(z), ptr(z), None, seed,
             it, row0, C, None, None, ptr(acc), ptr(accept), None, ptr(part),
             ctypes.byref(npart), None, stream())
    np.testing.assert_array_equal(N(accept), (u_ref < N(acc)).astype(np.int32))


@pytest.mark.parametrize("C,D", [(37, 24), (300, 1024), (5, 4), (64, 100)])
def test_momentum_vec4_path_is_the_scalar_path(C, D):
    """hmc.py:21-23.  The 128-bit momentum kernel (row_len % 4 == 0, aligned) and the scalar one
    (taken for a misaligned p) draw the same Philox blocks and do the same arithmetic: identical
    bits for p and for the kinetic energy; both match the oracle's Philox normals."""
    from zhusuan_b200._lib import lib, ptr, stream
    from oracle import philox
    g = np.random.RandomState(5)
    mass = torch.tensor(g.uniform(0.3, 3.0, D), dtype=torch.float32, device="cuda")
    seed, it, row0 = 0xFEEDFACE12345678, 77, 4096
    buf = torch.zeros(C * D + 4, device="cuda")
    p_al, p_mis = torch.empty(C, D, device="cuda"), buf[1:1 + C * D].view(C, D)
    assert p_al.data_ptr() % 16 == 0 and p_mis.data_ptr() % 16 != 0
    k_al = torch.empty(C, device="cuda")
    k_mis = torch.empty(C, device="cuda")
    for p, k in ((p_al, k_al), (p_mis, k_mis)):
        lib.call("zsb_hmc_momentum_f32", ptr(p), None, ptr(mass), D, C, D, seed, it, 1, row0,
                 ptr(k), 0, None, stream())
    np.testing.assert_array_equal(N(p_al), N(p_mis))
    np.testing.assert_array_equal(N(k_al), N(k_mis))
    ref = philox.normal_matrix(seed, 1, it, row0, C, D) * np.sqrt(N(mass))[None, :]
    np.testing.assert_allclose(N(p_al), ref, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(N(k_al), 0.5 * (ref.astype(np.float64) ** 2 / N(mass)).sum(1),
                               rtol=2e-5)
    # accumulate = 1 adds to the existing kinetic energy (second latent of a multi-latent model)
    k2 = k_al.clone()
    lib.call("zsb_hmc_momentum_f32", ptr(p_al), None, ptr(mass), D, C, D, seed, it, 1, row0,
             ptr(k2), 1, None, stream())
    np.testing.assert_allclose(N(k2), 2 * N(k_al), rtol=1e-6)


# ---- the nine other elementwise univariate families (univariate_ext.cu) ---------------------
def _uni_dist(zs, fam, a, b):
    D = zs.distributions
    t = lambda v: torch.tensor(np.asarray(v), dtype=torch.float32, device="cuda")
    if fam == "fold_normal":
        return D.FoldNormal(t(a), logstd=t(b))
    if fam == "uniform":
        return D.Uniform(t(a), t(b))
    if fam == "gamma":
        return D.Gamma(t(a), t(b))
    if fam == "inverse_gamma":
        return D.InverseGamma(t(a), t(b))
    if fam == "beta":
        return D.Beta(t(a), t(b))
    if fam == "poisson":
        return D.Poisson(t(a))
    if fam == "binomial":
        return D.Binomial(t(a), int(b))
    if fam == "laplace":
        return D.Laplace(t(a), t(b))
    return D.BinConcrete(t(a), t(b))


def test_univariate_more_reference_values():
    """tests/distributions/test_univariate.py `_test_value` literals (cases.py) through the
    Distribution classes: log_prob and prob against the scipy.stats targets."""
    import zhusuan_b200 as zs
    for fam, given, a, b, target, atol in cases.univariate_more_cases():
        d = _uni_dist(zs, fam, a, b)
        g = torch.tensor(given, device="cuda")
        with np.errstate(all="ignore"):
            lp = d.log_prob(g).cpu().numpy()
            p = d.prob(g).cpu().numpy()
            np.testing.assert_allclose(lp, target, rtol=2e-5, atol=max(atol, 1e-4), err_msg=fam)
            np.testing.assert_allclose(p, np.exp(target), rtol=1e-3, atol=1e-6, err_msg=fam)


@pytest.mark.parametrize("fam", ["fold_normal", "gamma", "inverse_gamma", "beta", "poisson",
                                 "binomial", "laplace", "bin_concrete", "uniform"])
def test_univariate_more_vs_oracle_with_gradients(fam):
    """Random broadcast shapes + group_ndims: forward vs the NumPy oracle, backward vs a
    float64 torch-autograd restatement of the same formula."""
    import zhusuan_b200 as zs
    from oracle import distributions as OD
    rng = np.random.RandomState(sum(map(ord, fam)))
    shape, pshape = (5, 7, 6), (7, 6)
    pos = lambda s: (0.3 + 3 * rng.random_sample(s)).astype(np.float32)
    unit = lambda s: (0.02 + 0.96 * rng.random_sample(s)).astype(np.float32)
    lgam = torch.lgamma
    sp = torch.nn.functional.softplus
    if fam == "fold_normal":
        x, a, b = pos(shape), rng.standard_normal(pshape).astype(np.float32), \
            (0.3 * rng.standard_normal((6,))).astype(np.float32)
        ref = lambda x, a, b: (-0.5 * np.log(2 * np.pi) - (b + 0.5 * torch.exp(-2 * b) * (x - a) ** 2)
                               + sp(-2 * a * x * torch.exp(-2 * b)))
    elif fam == "gamma":
        x, a, b = pos(shape), pos(pshape), pos((6,))
        ref = lambda x, a, b: a * torch.log(b) - lgam(a) + (a - 1) * torch.log(x) - b * x
    elif fam == "inverse_gamma":
        x, a, b = pos(shape), pos(pshape), pos((6,))
        ref = lambda x, a, b: a * torch.log(b) - lgam(a) - (a + 1) * torch.log(x) - b / x
    elif fam == "beta":
        x, a, b = unit(shape), pos(pshape), pos((6,))
        ref = lambda x, a, b: (a - 1) * torch.log(x) + (b - 1) * torch.log(1 - x) - (
            lgam(a) + lgam(b) - lgam(a + b))
    elif fam == "poisson":
        x, a, b = rng.poisson(3., shape).astype(np.float32), pos(pshape), None
        ref = lambda x, a, b: x * torch.log(a) - a - lgam(x + 1)
    elif fam == "binomial":
        x, a, b = rng.binomial(12, 0.4, shape).astype(np.float32), \
            rng.standard_normal(pshape).astype(np.float32), 12
        ref = lambda x, a, b: (lgam(b + 1) - lgam(b - x + 1) - lgam(x + 1) + x * a - b * sp(a))
    elif fam == "laplace":
        x, a, b = rng.standard_normal(shape).astype(np.float32), \
            rng.standard_normal(pshape).astype(np.float32), pos((6,))
        ref = lambda x, a, b: -np.log(2.) - torch.log(b) - torch.abs(x - a) / b
    elif fam == "bin_concrete":
        x, a, b = unit(shape), np.float32(0.7), rng.standard_normal(pshape).astype(np.float32)

        def ref(x, a, b):
            t = a * (torch.log(x) - torch.log(1 - x)) - b
            return torch.log(a) - torch.log(x) - torch.log(1 - x) + t - 2 * sp(t)
    else:  # uniform
        a, b = (-1 - rng.random_sample(pshape)).astype(np.float32), pos((6,))
        x = (rng.random_sample(shape) * 0.9 - 0.5).astype(np.float32)
        ref = lambda x, a, b: -torch.log(b - a) + 0 * x
    for gnd in (0, 2):
        tx = torch.tensor(x, device="cuda", requires_grad=(fam not in ("poisson", "binomial")))
        ta = torch.tensor(a, device="cuda", requires_grad=True)
        tb = torch.tensor(b, device="cuda", requires_grad=True) \
            if (b is not None and fam != "binomial") else None
        if fam == "fold_normal":
            d = zs.distributions.FoldNormal(ta, logstd=tb, group_ndims=gnd)
        elif fam == "poisson":
            d = zs.distributions.Poisson(ta, group_ndims=gnd)
        elif fam == "binomial":
            d = zs.distributions.Binomial(ta, 12, group_ndims=gnd)
        else:
            cls = dict(gamma="Gamma", inverse_gamma="InverseGamma", beta="Beta", laplace="Laplace",
                       bin_concrete="BinConcrete", uniform="Uniform")[fam]
            d = getattr(zs.distributions, cls)(ta, tb, group_ndims=gnd)
        lp = d.log_prob(tx)
        ofn = getattr(OD, fam + "_log_prob")
        want = ofn(x, a, b, group_ndims=gnd, dtype=np.float64) if b is not None else \
            ofn(x, a, group_ndims=gnd, dtype=np.float64)
        np.testing.assert_allclose(lp.detach().cpu().numpy(), want, rtol=2e-5, atol=2e-5)
        w = torch.tensor(rng.standard_normal(tuple(lp.shape)), dtype=torch.float32, device="cuda")
        ins = [t for t in (tx, ta, tb) if t is not None and t.requires_grad]
        got = torch.autograd.grad((lp * w).sum(), ins)
        # float64 reference on the host
        rx = torch.tensor(x, dtype=torch.float64, requires_grad=tx.requires_grad)
        ra = torch.tensor(a, dtype=torch.float64, requires_grad=True)
        rb = torch.tensor(b, dtype=torch.float64, requires_grad=tb is not None) \
            if b is not None else None
        rl = ref(rx, ra, rb)
        if gnd:
            rl = rl.sum(dim=tuple(range(-gnd, 0)))
        rins = [t for t in (rx, ra, rb) if t is not None and t.requires_grad]
        exp = torch.autograd.grad((rl * w.double().cpu()).sum(), rins)
        for gg, ee in zip(got, exp):
            np.testing.assert_allclose(gg.cpu().numpy(), ee.numpy(), rtol=3e-4, atol=3e-4)


def test_univariate_more_contract_and_sampling():
    """Constructor checks (messages of the reference) and sample shapes / supports / moments."""
    import zhusuan_b200 as zs
    D = zs.distributions
    dev = "cuda"
    one = torch.ones(3, device=dev)
    with pytest.raises(ValueError, match="should be broadcastable to match"):
        D.Gamma(torch.ones(2, device=dev), one)
    with pytest.raises(ValueError, match="Either std or logstd"):
        D.FoldNormal(one)
    with pytest.raises(ValueError, match="n_experiments must be positive"):
        D.Binomial(one, 0)
    with pytest.raises(TypeError, match="n_experiments must be int32"):
        D.Binomial(one, 2.5)
    with pytest.raises(TypeError, match="must have the same dtype as"):
        D.Laplace(one, one.double())
    with pytest.raises(ValueError, match="should be a scalar"):
        D.BinConcrete(one, one)
    n = 20000
    a, b = torch.tensor([2.0, 5.0], device=dev), torch.tensor([1.5, 2.5], device=dev)
    for d, mean in [(D.Gamma(a, b), a / b), (D.Beta(a, b), a / (a + b)),
                    (D.InverseGamma(a + 2, b), b / (a + 1)), (D.Poisson(a), a),
                    (D.Binomial(torch.zeros(2, device=dev), 10), torch.full((2,), 5.0)),
                    (D.Laplace(a, b), a), (D.Uniform(a, a + b), a + b / 2)]:
        s = d.sample(n)
        assert tuple(s.shape) == (n, 2) and s.dtype == d.dtype
        np.testing.assert_allclose(s.float().mean(0).cpu().numpy(), mean.cpu().numpy(), rtol=0.08)
        assert torch.isfinite(d.log_prob(s)).all()
    s = D.BinConcrete(torch.tensor(0.5, device=dev), torch.zeros(4, device=dev)).sample(100)
    # sigmoid((logit + logistic noise) / T) saturates to exactly 0 / 1 in float32 for |noise| > ~8 T
    assert tuple(s.shape) == (100, 4) and bool(((s >= 0) & (s <= 1)).all())
    # BayesianNet factories
    bn = zs.BayesianNet(observed={"g": torch.tensor([1.0, 2.0], device=dev)})
    g = bn.gamma("g", a, b, group_ndims=1)
    bn.laplace("l", a, b, n_samples=5)
    bn.poisson("k", a)
    assert tuple(g.cond_log_p.shape) == () and tuple(bn["l"].tensor.shape) == (5, 2)
    assert tuple(bn.log_joint().shape) == (5, 2)


def test_multinomial_and_onehot_categorical():
    """tests/distributions/test_multivariate.py:218-253, 505-533 through the classes (log_prob on
    the unnormalised-multinomial kernel), gradients wrt logits, sampling, factories."""
    import zhusuan_b200 as zs
    D = zs.distributions
    for l, n, g, normalize, tgt in cases.multinomial_cases():
        for ne in (None, n):
            d = D.Multinomial(torch.tensor(l, device="cuda"), ne, normalize_logits=normalize)
            lp = d.log_prob(torch.tensor(g, device="cuda"))
            np.testing.assert_allclose(lp.cpu().numpy(), tgt, rtol=1e-5, atol=1e-4)
    rng = np.random.RandomState(0)
    logits = rng.standard_normal((4, 6)).astype(np.float32)
    tl = torch.tensor(logits, device="cuda", requires_grad=True)
    counts = rng.multinomial(9, np.ones(6) / 6, size=(3, 4)).astype(np.int32)
    d = D.Multinomial(tl, 9, group_ndims=1)
    lp = d.log_prob(torch.tensor(counts, device="cuda"))
    np.testing.assert_allclose(lp.detach().cpu().numpy(),
                               OD.multinomial_log_prob(counts, logits, 9, group_ndims=1,
                                                       dtype=np.float64), rtol=1e-5, atol=1e-4)
    g, = torch.autograd.grad(lp.sum(), [tl])
    soft = np.exp(logits - logits.max(-1, keepdims=True)); soft /= soft.sum(-1, keepdims=True)
    np.testing.assert_allclose(g.cpu().numpy(), (counts - 9 * soft).sum(0), rtol=1e-4, atol=1e-4)
    s = d.sample(5)
    assert tuple(s.shape) == (5, 4, 6) and s.dtype == torch.int32
    assert bool((s.sum(-1) == 9).all())
    with pytest.raises(ValueError, match="Cannot sample"):
        D.Multinomial(tl, None).sample(1)
    oh = D.OnehotCategorical(tl)
    so = oh.sample(7)
    assert tuple(so.shape) == (7, 4, 6) and bool((so.sum(-1) == 1).all())
    idx = so.argmax(-1)
    np.testing.assert_allclose(oh.log_prob(so).detach().cpu().numpy(),
                               D.Categorical(tl).log_prob(idx).detach().cpu().numpy(), rtol=1e-5,
                               atol=1e-6)
    bn = zs.BayesianNet()
    bn.multinomial("m", tl, 9, n_samples=2)
    bn.onehot_categorical("o", tl)
    assert tuple(bn["m"].tensor.shape) == (2, 4, 6) and tuple(bn["o"].tensor.shape) == (4, 6)


def test_concrete_family_and_matrix_variate_normal():
    """ExpConcrete / Concrete (multivariate.py:683-958) and MatrixVariateNormalCholesky
    (:961-1160) through the classes: values vs the oracle / scipy, reparameterised sampling."""
    import zhusuan_b200 as zs
    from scipy import stats
    D = zs.distributions
    rng = np.random.RandomState(4)
    logits = rng.standard_normal((5, 7)).astype(np.float32)
    tl = torch.tensor(logits, device="cuda", requires_grad=True)
    tt = torch.tensor(0.7, device="cuda", requires_grad=True)
    for cls, ofn in ((D.Concrete, OD.concrete_log_prob), (D.ExpConcrete, OD.exp_concrete_log_prob)):
        d = cls(tt, tl, group_ndims=0)
        s = d.sample(6)
        assert tuple(s.shape) == (6, 5, 7)
        tot = s.sum(-1) if cls is D.Concrete else torch.exp(s).sum(-1)
        np.testing.assert_allclose(tot.detach().cpu().numpy(), 1.0, rtol=1e-4)
        lp = d.log_prob(s.detach())
        want = ofn(s.detach().cpu().numpy(), 0.7, logits, dtype=np.float64)
        np.testing.assert_allclose(lp.detach().cpu().numpy(), want, rtol=1e-4, atol=1e-3)
        g = torch.autograd.grad(lp.sum() + s.sum() * 0, [tl, tt], allow_unused=True)
        assert g[0] is not None and torch.isfinite(g[0]).all()
        assert tuple(cls(tt, tl, group_ndims=1).log_prob(s.detach()).shape) == (6,)
    with pytest.raises(ValueError, match="should be a scalar"):
        D.Concrete(torch.ones(2, device="cuda"), tl)
    # matrix-variate normal, batch [2]
    r, c = 3, 4
    us, vs, lus, lvs = [], [], [], []
    for _ in range(2):
        a = rng.standard_normal((r, r)); u = a @ a.T + r * np.eye(r)
        b = rng.standard_normal((c, c)); v = b @ b.T + c * np.eye(c)
        us.append(u); vs.append(v)
        lus.append(np.linalg.cholesky(u)); lvs.append(np.linalg.cholesky(v))
    mean = rng.standard_normal((2, r, c)).astype(np.float32)
    f = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")
    tm = f(mean).requires_grad_(True)
    dst = D.MatrixVariateNormalCholesky(tm, f(lus), f(lvs))
    x = dst.sample(50)
    assert tuple(x.shape) == (50, 2, r, c) and x.requires_grad
    lp = dst.log_prob(x.detach())
    assert tuple(lp.shape) == (50, 2)
    for b in range(2):
        want = stats.matrix_normal.logpdf(x[:, b].detach().cpu().numpy().astype(np.float64),
                                          mean[b], us[b], vs[b])
        np.testing.assert_allclose(lp[:, b].detach().cpu().numpy(), want, rtol=1e-4, atol=1e-3)
    bn = zs.BayesianNet()
    bn.concrete("c", tt, tl, n_samples=2)
    bn.matrix_variate_normal_cholesky("m", tm, f(lus), f(lvs))
    assert tuple(bn["c"].tensor.shape) == (2, 5, 7) and tuple(bn["m"].tensor.shape) == (2, r, c)


def test_plugin_distribution_subclass_group_sum_on_device():
    """A user-defined Distribution (the plugin contract, tests/distributions/test_base.py:15-140)
    returning the un-grouped log density gets the `group_ndims` sum from the base class, through
    the group-sum kernel, and works as a BayesianNet node."""
    import zhusuan_b200 as zs

    class Scaled(zs.distributions.Distribution):
        def __init__(self, scale, group_ndims=0):
            self.scale = scale
            super(Scaled, self).__init__(torch.float32, torch.float32, is_continuous=True,
                                         is_reparameterized=True, group_ndims=group_ndims)

        def _get_value_shape(self):
            return torch.Size([])

        def _get_batch_shape(self):
            return self.scale.shape

        def _sample(self, n_samples):
            return torch.ones((n_samples,) + tuple(self.scale.shape), device=self.scale.device)

        def _log_prob(self, given):
            return -given * self.scale

    scale = torch.arange(1., 25., device="cuda").reshape(2, 3, 4).requires_grad_(True)
    x = torch.ones(5, 2, 3, 4, device="cuda")
    for g, shape in ((0, (5, 2, 3, 4)), (1, (5, 2, 3)), (3, (5,))):
        lp = Scaled(scale, group_ndims=g).log_prob(x)
        assert tuple(lp.shape) == shape
        want = -(x * scale).detach()
        if g:
            want = want.sum(dim=tuple(range(-g, 0)))
        np.testing.assert_allclose(lp.detach().cpu().numpy(), want.cpu().numpy(), rtol=1e-6)
    gr, = torch.autograd.grad(Scaled(scale, group_ndims=2).log_prob(x).sum(), [scale])
    np.testing.assert_allclose(gr.cpu().numpy(), -5.0)
    bn = zs.BayesianNet(observed={"y": x})
    node = bn.stochastic("y", Scaled(scale, group_ndims=3))
    assert tuple(node.cond_log_p.shape) == (5,) and tuple(bn.log_joint().shape) == (5,)


@pytest.mark.parametrize("rows,group,pn", [(4096, 16, "row"), (1000, 40, "tile"), (64, 784, "full"),
                                           (333, 12, "scalar"), (50, 100, "row")])
def test_vec4_row_kernels_match_scalar_path(rows, group, pn):
    """csrc/distributions.cu: the 128-bit row-reduce path (aligned operands, group % 4 == 0) and
    the scalar path (forced here by operands whose device pointer is 4 bytes off a 16-byte
    boundary) evaluate the same log-probs, gradients and reparameterised samples -- the latter
    bit for bit, Philox block i >> 2 feeding elements 4 (i >> 2) .. + 3 in both."""
    from zhusuan_b200._lib import lib, ptr, stream
    rng = np.random.RandomState(rows + group)
    n = rows * group
    pshape = {"row": (group,), "tile": (rows // 8 if rows % 8 == 0 else rows, group),
              "full": (rows, group), "scalar": (1,)}[pn]

    def both(a):
        """(aligned copy, copy whose data_ptr is 4 bytes past a 16-byte boundary)"""
        al = T(a).contiguous()
        buf = torch.empty(al.numel() + 4, device="cuda", dtype=al.dtype)
        off = buf[1:1 + al.numel()].view(al.shape)
        off.copy_(al)
        assert al.data_ptr() % 16 == 0 and off.data_ptr() % 16 == 4
        return al, off
    x = both(rng.standard_normal((rows, group)).astype(np.float32))
    mu = both(rng.standard_normal(pshape).astype(np.float32))
    ls = both((0.3 * rng.standard_normal(pshape)).astype(np.float32))
    lg = both(rng.standard_normal((rows, group)).astype(np.float32))
    xb = both((rng.random_sample(pshape if pn != "scalar" else (rows, group)) < 0.3)
              .astype(np.float32))
    gout = T(rng.standard_normal(rows).astype(np.float32))
    res = []
    for v in (0, 1):
        out = torch.empty(rows, device="cuda")
        lib.call("zsb_logprob_normal_f32", ptr(x[v]), n, ptr(mu[v]), mu[v].numel(), ptr(ls[v]),
                 ls[v].numel(), ptr(out), rows, group, stream())
        d = [both(np.zeros((rows, group), np.float32))[v] for _ in range(3)]
        lib.call("zsb_logprob_normal_bwd_f32", ptr(x[v]), n, ptr(mu[v]), mu[v].numel(),
                 ptr(ls[v]), ls[v].numel(), ptr(gout), rows, group, ptr(d[0]), ptr(d[1]),
                 ptr(d[2]), stream())
        ob = torch.empty(rows, device="cuda")
        lib.call("zsb_logprob_bernoulli_f32", ptr(xb[v]), xb[v].numel(), ptr(lg[v]), n, ptr(ob),
                 rows, group, stream())
        dl = both(np.zeros((rows, group), np.float32))[v]
        lib.call("zsb_logprob_bernoulli_bwd_f32", ptr(xb[v]), xb[v].numel(), ptr(lg[v]), n,
                 ptr(gout), rows, group, ptr(dl), stream())
        gs = torch.empty(rows, device="cuda")
        lib.call("zsb_group_sum_f32", ptr(x[v]), ptr(gs), rows, group, stream())
        z = both(np.zeros((rows, group), np.float32))[v]
        e = both(np.zeros((rows, group), np.float32))[v]
        lq = torch.empty(rows, device="cuda")
        lib.call("zsb_reparam_normal_f32", ptr(mu[v]), mu[v].numel(), ptr(ls[v]), ls[v].numel(),
                 None, 77, 3, ptr(z), ptr(e), ptr(lq), rows, group, stream())
        z1 = both(np.zeros((rows, group), np.float32))[v]
        lib.call("zsb_reparam_normal_f32", ptr(mu[v]), mu[v].numel(), ptr(ls[v]), ls[v].numel(),
                 None, 77, 3, ptr(z1), None, None, n, 1, stream())
        res.append([N(t) for t in (out, d[0], d[1], d[2], ob, dl, gs, z, e, lq, z1)])
    names = "lp dgiven dmean dlogstd lp_b dlogits gsum z eps logq z_flat".split()
    for name, a, b in zip(names, *res):
        if name in ("z", "eps", "z_flat"):
            np.testing.assert_array_equal(a, b, err_msg=name)
        else:
            np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5 * group ** 0.5, err_msg=name)
    np.testing.assert_array_equal(res[0][7], res[0][10])       # grouped and flat draws agree
    ref = OD.normal_log_prob(N(x[0]), N(mu[0]).reshape(pshape), N(ls[0]).reshape(pshape), 1,
                             np.float64) if pn != "tile" else None
    if ref is not None:
        np.testing.assert_allclose(res[0][0], ref, rtol=RTOL, atol=1e-4)
