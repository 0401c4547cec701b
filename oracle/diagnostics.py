"""NumPy restatement of the effective-sample-size estimator (TEST ORACLE ONLY).

zhusuan/diagnostics.py:17-64 (the Stan estimator).  Vectorised over lags with an
FFT-free cumulative formulation: autocovariances are evaluated lag by lag until
the first negative autocorrelation, as the reference's loop does.
"""
import numpy as np


def ess_1d(chain):
    """diagnostics.py:17-41."""
    x = np.asarray(chain, np.float64)
    n = x.shape[0]
    c = x - x.mean()
    var_plus = np.mean(c * c)                    # np.var
    var = var_plus * n / (n - 1)
    total = 0.0
    for lag in range(n):
        acov = np.dot(c[:n - lag], c[lag:]) / (n - lag)
        rho = 1.0 - (var - acov) / var_plus
        if rho < 0:
            break
        total += rho
    return n / (1.0 + 2.0 * total)


def ess_per_dim(samples, burn_in=100):
    s = np.asarray(samples)[burn_in:]
    return np.array([ess_1d(s[:, d]) for d in range(s.shape[1])])


def ess(samples, burn_in=100):
    """diagnostics.py:44-64: min over the positive per-dimension values."""
    e = ess_per_dim(samples, burn_in)
    assert (e >= 0).all()
    pos = e[e > 0]
    return pos.min() if pos.size else np.inf
