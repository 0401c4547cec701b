"""NumPy restatement of zhusuan/evaluation.py:57-172 (AIS) on the HMC oracle
(TEST ORACLE ONLY -- see oracle/__init__).  Every random draw is injected: the two prior
samples (evaluation.py:87, 100-101, 138) and the HMC noise of each adaptation / temperature
iteration.  Pinned to the reference's own class AIS run on the NumPy TF stand-in
(oracle/tf_shim/make_ref_golden.py -> tests/golden/ref_ais.npz, tests/test_ref_pins.py);
tests/test_gpu_models.py also re-runs the reference's statistical check (known marginal) and
compares the device loop with this restatement and with that fixture step for step."""
import numpy as np

from . import hmc as OH


class AIS(object):
    def __init__(self, log_prior, grad_prior, log_joint, grad_joint, hmc, n_temperatures=1000,
                 n_adapt=30, dtype=np.float32):
        """log_* / grad_*: callables on a list of latent arrays (as oracle.hmc.HMC takes);
        hmc: an oracle.hmc.HMC instance."""
        self.lp, self.gp, self.lj, self.gj = log_prior, grad_prior, log_joint, grad_joint
        self.hmc = hmc
        self.n_temperatures = int(n_temperatures)
        self.n_adapt = int(n_adapt)
        self.dtype = dtype
        self.temperature = dtype(0)

    def _map_t(self, t):                                    # evaluation.py:107-108
        return 1. / (1. + np.exp(-4 * (2 * t / self.n_temperatures - 1)))

    def schedule(self, t):                                  # evaluation.py:110-112
        return (self._map_t(t) - self._map_t(0)) / (
            self._map_t(self.n_temperatures) - self._map_t(0))

    def _logp(self, q):                                     # evaluation.py:91-94
        d, t = self.dtype, self.temperature
        return (self.lp(q).astype(d) * (d(1) - t) + self.lj(q).astype(d) * t).astype(d)

    def _grad(self, q):
        d, t = self.dtype, self.temperature
        return [(a.astype(d) * (d(1) - t) + b.astype(d) * t).astype(d)
                for a, b in zip(self.gp(q), self.gj(q))]

    def run(self, init, noise, adapt_flags=(True, True)):
        """init: the two prior draws (lists of arrays); noise(k) -> (noise_p list, noise_u).
        Returns (bound, log_weights)."""
        d = self.dtype
        step = 0
        adp = 2 if self.n_temperatures > 1 else 1
        q = [np.asarray(x, d) for x in init[0]]
        self.temperature = d(self.schedule(adp))
        for _ in range(self.n_adapt):                       # evaluation.py:129-136
            npz, nu = noise(step); step += 1
            q, _ = self.hmc.step(q, self._logp, self._grad, npz, nu, *adapt_flags)
        q = [np.asarray(x, d) for x in init[1]]             # evaluation.py:138
        self.temperature = d(0)
        log_w = -self._logp(q)                              # evaluation.py:139-143
        for num_t in range(self.n_temperatures):
            self.temperature = d(self.schedule(num_t + 1))
            npz, nu = noise(step); step += 1
            q, info = self.hmc.step(q, self._logp, self._grad, npz, nu, *adapt_flags)
            if num_t + 1 < self.n_temperatures:             # evaluation.py:155-158
                log_w = (log_w + (info.orig_log_prob - info.log_prob)).astype(d)
            else:
                log_w = (log_w + info.orig_log_prob).astype(d)
        m = log_w.max(0)                                    # evaluation.py:167-172
        bound = np.log(np.mean(np.exp(log_w - m), 0)) + m
        return float(np.mean(bound)), log_w
