"""Oracle pin: effective sample size against the reference's own test
(tests/test_diagnostics.py:14-42): i.i.d. Gaussian draws keep ESS >= 2000 of
10 000, a random-walk Metropolis chain drops below 1000."""
import numpy as np

from oracle import diagnostics as OG


def make_chains():
    rng = np.random.RandomState(1)
    n, dims = 10000, 2
    iid = rng.normal(size=(n, dims))
    cur = np.zeros(dims)
    acc_rng = np.random.RandomState(2)      # the reference uses the global np.random here
    mcmc = []
    for _ in range(n):
        nxt = cur + rng.normal(size=dims)
        a = np.exp(np.minimum(0, -0.5 * np.sum(nxt ** 2 - cur ** 2)))
        if acc_rng.random_sample() < a:
            cur = nxt
        mcmc.append(cur.copy())
    return iid, np.array(mcmc)


def test_reference_properties():
    iid, mcmc = make_chains()
    assert OG.ess(iid, burn_in=100) >= 2000
    assert OG.ess(mcmc, burn_in=100) <= 1000


def test_definition_small_case():
    x = np.array([0.3, -1.2, 0.8, 0.5, -0.1, 1.4, -0.7, 0.2])
    n = len(x)
    mu, vp = x.mean(), x.var()
    v = vp * n / (n - 1)
    tot = 0.0
    for t in range(n):
        ac = np.mean((x[:n - t] - mu) * (x[t:] - mu))
        rho = 1 - (v - ac) / vp
        if rho < 0:
            break
        tot += rho
    assert abs(OG.ess_1d(x) - n / (1 + 2 * tot)) < 1e-12
