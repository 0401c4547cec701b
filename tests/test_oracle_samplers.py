"""CPU checks of oracle/samplers.py (the restatement of csrc/samplers.cu): exact distribution
moments / frequencies (what the reference's own sampler tests check, tests/distributions/utils.py)
and the inverse-CDF edge cases."""
import numpy as np

from oracle import samplers as OS


def test_gamma_marsaglia_tsang_moments():
    a = np.array([0.4, 1.0, 2.5, 9.0])
    g = OS.gamma_marsaglia_tsang(np.tile(a, 30000), 123, 5).reshape(-1, 4)
    assert (g > 0).all()
    np.testing.assert_allclose(g.mean(0), a, rtol=0.03)
    np.testing.assert_allclose(g.var(0), a, rtol=0.08)


def test_dirichlet_moments_and_simplex():
    alpha = np.array([[0.5, 1.5, 3.0], [2.0, 2.0, 2.0]])
    x = OS.dirichlet(alpha, 20000, 9, 2)
    np.testing.assert_allclose(x.sum(-1), 1.0, rtol=1e-12)
    mean = alpha / alpha.sum(-1, keepdims=True)
    np.testing.assert_allclose(x.mean(0), mean, atol=0.01)


def test_categorical_inverse_cdf_edges_and_frequencies():
    l = np.log(np.array([[0.2, 0.3, 0.5], [1e-30, 0.5, 0.5]]))
    l[1, 0] = -np.inf
    u = np.stack([np.zeros(2), np.full(2, 1 - 2.0 ** -24), np.full(2, 0.2), np.full(2, 0.5)])
    idx = OS.categorical_inverse_cdf(l, u)
    assert idx[0].tolist() == [0, 1]          # u = 0 -> first category with mass
    assert idx[1].tolist() == [2, 2]          # u -> 1 -> last category
    assert idx[2, 0] == 1                     # cdf 0.2 is not > 0.2
    assert idx[3, 1] == 2                     # cdf 0.5 is not > 0.5
    uu = OS.categorical_uniforms(7, 3, 60000).reshape(-1, 2)
    assert uu.min() >= 0 and uu.max() < 1
    out = OS.categorical_inverse_cdf(l, uu)
    np.testing.assert_allclose([(out[:, 0] == c).mean() for c in range(3)], [0.2, 0.3, 0.5],
                               atol=0.01)
    assert (out[:, 1] != 0).all()
