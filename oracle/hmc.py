"""NumPy restatement of zhusuan/hmc.py (TEST ORACLE ONLY -- see oracle/__init__).

Every random draw of the reference (``tf.random_normal`` hmc.py:22,
``tf.random_uniform`` hmc.py:485) is an INJECTED argument, because TF's Philox
streams cannot be reproduced without TensorFlow.  The arithmetic type is
``dtype`` (float32 mirrors hmc.py's hard-coded tf.float32; float64 is the
twin used to bound rounding error).

The target is given as two callables on a list of latent arrays
(``chain axes + data axes`` each):
    logp(q_list) -> array[chain axes]        (hmc.py:426-428)
    grad(q_list) -> list of arrays like q    (hmc.py:430-432, tf.gradients)
"""
import numpy as np


def _f(dtype, x):
    return dtype(x)


class StepsizeTuner(object):
    """hmc.py:64-112 (Nesterov dual averaging; note mu = 10*eps0, line 79)."""

    def __init__(self, initial_stepsize, gamma, t0, kappa, delta,
                 dtype=np.float32):
        d = dtype
        self.dtype = d
        self.gamma, self.t0, self.kappa, self.delta = (
            d(gamma), d(t0), d(kappa), d(delta))
        self.mu = d(10 * initial_stepsize)          # hmc.py:79
        self.step = d(0)                            # hmc.py:82
        self.log_epsilon_bar = d(0)                 # hmc.py:84
        self.h_bar = d(0)                           # hmc.py:86

    def tune(self, adapt, acceptance_rate, fresh_start):
        d = self.dtype
        one = d(1)
        if adapt:                                   # hmc.py:91-106
            fresh_start = d(fresh_start)
            self.step = (one - fresh_start) * self.step + one
            rate1 = one / (self.step + self.t0)
            self.h_bar = ((one - fresh_start) * (one - rate1) * self.h_bar
                          + rate1 * (self.delta - d(acceptance_rate)))
            log_epsilon = self.mu - np.sqrt(self.step) / self.gamma * \
                self.h_bar
            rate = np.power(self.step, -self.kappa)
            self.log_epsilon_bar = (
                rate * log_epsilon
                + (one - fresh_start) * (one - rate) * self.log_epsilon_bar)
            return d(np.exp(log_epsilon))
        return d(np.exp(self.log_epsilon_bar))      # hmc.py:110


class ExponentialWeightedMovingVariance(object):
    """hmc.py:115-159; mean/var have shape [1..1, data dims]."""

    def __init__(self, decay, shapes, num_chain_dims, dtype=np.float32):
        self.dtype = dtype
        self.t = dtype(0)
        self.mean = [np.zeros(s, dtype) for s in shapes]
        self.var = [np.zeros(s, dtype) for s in shapes]
        self.decay = dtype(decay)
        self.chain_axes = tuple(range(num_chain_dims))

    def update(self, x):
        d = self.dtype
        self.t = self.t + d(1)                                   # :132
        weight = (d(1) - self.decay) / (d(1) - np.power(self.decay, self.t))
        weight = d(weight)
        new_var = []
        for k, q in enumerate(x):
            incr = weight * (q - self.mean[k])                   # :135
            self.mean[k] = self.mean[k] + incr.mean(
                axis=self.chain_axes, keepdims=True, dtype=d)    # :137-139
            nv = (d(1) - weight) * self.var[k] + (
                incr * (q - self.mean[k])).mean(
                    axis=self.chain_axes, keepdims=True, dtype=d)  # :141-145
            new_var.append(nv.astype(d))
        self.var = new_var
        return self.var

    def precision(self):
        with np.errstate(divide='ignore'):
            return [(self.dtype(1) / v).astype(self.dtype) for v in self.var]


class HMCInfo(object):
    """hmc.py:162-201."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class HMC(object):
    """hmc.py:204-522, one ``step()`` == one ``sess.run(sample_op)``.

    ``adapt_step_size`` / ``adapt_mass`` given to the constructor only say
    whether the feature is configured (``is not None`` in the reference); the
    per-iteration boolean (the fed placeholder) is passed to ``step``.
    """

    def __init__(self, step_size=1., n_leapfrogs=10, adapt_step_size=None,
                 target_acceptance_rate=0.8, gamma=0.05, t0=100, kappa=0.75,
                 adapt_mass=None, mass_collect_iters=10, mass_decay=0.99,
                 dtype=np.float32):
        d = dtype
        self.dtype = d
        self.step_size = d(step_size)                       # hmc.py:258
        self.n_leapfrogs = int(n_leapfrogs)
        self.target_acceptance_rate = d(target_acceptance_rate)
        self.t = d(0)                                       # hmc.py:264
        self.has_step_adapt = adapt_step_size is not None
        if self.has_step_adapt:
            self.tuner = StepsizeTuner(step_size, gamma, t0, kappa,
                                       target_acceptance_rate, d)
        self.has_mass_adapt = adapt_mass is not None
        if self.has_mass_adapt:
            if not self.has_step_adapt:                     # hmc.py:271-272
                raise ValueError('If adapt mass is set, we should also adapt '
                                 'step size')
        else:
            mass_collect_iters = 0                          # hmc.py:276
        self.mass_collect_iters = int(mass_collect_iters)
        self.mass_decay = d(mass_decay)
        self.ewmv = None
        self.n_search_iters = 0     # diagnostics: passes of the search loop

    # -- helpers ---------------------------------------------------------
    def _kinetic(self, p, mass, data_axes):
        d = self.dtype
        k = None
        for pm, m, ax in zip(p, mass, data_axes):           # hmc.py:32-34
            term = (np.square(pm) / m).sum(axis=tuple(ax), dtype=d)
            k = term if k is None else k + term
        return (d(0.5) * k).astype(d)

    def _hamiltonian(self, q, p, logp, mass, data_axes):
        lp = logp(q).astype(self.dtype)
        return (-lp + self._kinetic(p, mass, data_axes)).astype(self.dtype), lp

    def _leapfrog_integrator(self, q, p, s1, s2, grad, mass):
        d = self.dtype
        q = [(x + d(s1) * (y / m)).astype(d)
             for x, y, m in zip(q, p, mass)]                # hmc.py:39, 26-27
        g = grad(q)                                         # hmc.py:41
        p = [(x + d(s2) * y.astype(d)).astype(d) for x, y in zip(p, g)]
        return q, p

    def _acceptance(self, q, p, nq, np_, logp, mass, data_axes):
        d = self.dtype
        h0, lp0 = self._hamiltonian(q, p, logp, mass, data_axes)
        h1, lp1 = self._hamiltonian(nq, np_, logp, mass, data_axes)
        if not np.all(np.isfinite(lp0)):                    # hmc.py:51-53
            raise FloatingPointError(
                'HMC: old_log_prob has numeric errors! Try better '
                'initialization.')
        with np.errstate(over='ignore', invalid='ignore'):
            acc = np.exp(np.minimum(-h1 + h0, d(0))).astype(d)  # :54-55
        ok = np.isfinite(acc) & np.isfinite(lp1)            # hmc.py:56-57
        acc = np.where(ok, acc, d(0)).astype(d)             # hmc.py:58-59
        return h0, h1, lp0, lp1, acc

    def _init_step_size(self, q, p, mass, grad, logp, data_axes):
        """hmc.py:307-345 (factor 1.5 search loop)."""
        d = self.dtype
        factor = d(1.5)
        step_size, last, cond = self.step_size, d(1.0), True
        while cond:
            self.n_search_iters += 1
            nq, np_ = self._leapfrog_integrator(
                q, p, d(0), step_size / d(2), grad, mass)
            nq, np_ = self._leapfrog_integrator(
                nq, np_, step_size, step_size / d(2), grad, mass)
            acc = self._acceptance(q, p, nq, np_, logp, mass, data_axes)[4]
            a = d(acc.mean(dtype=d))
            if a < self.target_acceptance_rate:             # hmc.py:329-333
                new_step = d(step_size * (d(1.0) / factor))
            else:
                new_step = d(step_size * factor)
            cond = not ((last < self.target_acceptance_rate)
                        ^ (a < self.target_acceptance_rate))  # hmc.py:335-337
            step_size, last = new_step, a
        return step_size

    # -- one iteration ---------------------------------------------------
    def step(self, q, logp, grad, noise_p, noise_u, adapt_step_size=False,
             adapt_mass=False):
        """One HMC iteration (hmc.py:382-522).

        q: list of latent arrays (updated copies are returned in info.samples).
        noise_p: list of standard-normal arrays like q (hmc.py:22).
        noise_u: uniform[0,1) array of chain shape (hmc.py:485).
        Returns (new_q list, HMCInfo).
        """
        d = self.dtype
        q = [np.asarray(x, d) for x in q]
        new_t = self.t + d(1)                               # hmc.py:418
        self.t = new_t
        chain_shape = logp(q).shape                         # hmc.py:436
        if len(chain_shape) == 0:                           # hmc.py:438-442
            raise ValueError('HMC requires that the static shape of the '
                             'value returned by log joint function should be '
                             'at least partially defined.')
        ncd = len(chain_shape)
        data_shapes = [(1,) * ncd + x.shape[ncd:] for x in q]   # :445-447
        data_axes = [list(range(ncd, x.ndim)) for x in q]       # :448-449

        # mass (hmc.py:452-456, 283-305)
        if self.has_mass_adapt:
            if self.ewmv is None:
                self.ewmv = ExponentialWeightedMovingVariance(
                    self.mass_decay, data_shapes, ncd, d)
            if adapt_mass:
                self.ewmv.update(q)
            new_mass = self.ewmv.precision()
            if int(new_t) < self.mass_collect_iters:        # hmc.py:299-302
                mass = [np.ones(s, d) for s in data_shapes]
            else:
                mass = new_mass
        else:
            mass = [np.ones(s, d) for s in data_shapes]

        # momentum (hmc.py:458, 21-23)
        p = [(np.asarray(n, d) * np.sqrt(m)).astype(d)
             for n, m in zip(noise_p, mass)]

        # step size for this iteration (hmc.py:463-472)
        if not self.has_step_adapt:
            eps = self.step_size
            init = False
        else:
            init = (new_t == d(1)) or (int(new_t) == self.mass_collect_iters)
            if init:
                eps = self._init_step_size(q, p, mass, grad, logp, data_axes)
            else:
                eps = self.step_size

        # leapfrog (hmc.py:347-372)
        cq, cp = q, p
        L = self.n_leapfrogs
        for i in range(L + 1):
            s1 = eps if i > 0 else d(0)
            s2 = eps if (0 < i < L) else eps / d(2)
            cq, cp = self._leapfrog_integrator(cq, cp, s1, s2, grad, mass)

        # MH test (hmc.py:479-498)
        h0, h1, lp0, lp1, acc = self._acceptance(q, p, cq, cp, logp, mass,
                                                 data_axes)
        u = np.asarray(noise_u, d)
        if_accept = u < acc
        new_q = []
        for nq, oq, da in zip(cq, q, data_axes):
            e = if_accept.reshape(if_accept.shape + (1,) * len(da))
            new_q.append(np.where(e, nq, oq).astype(d))
        new_lp = np.where(if_accept, lp1, lp0).astype(d)

        # step-size adaptation (hmc.py:501-505, 374-380)
        if self.has_step_adapt:
            self.step_size = self.tuner.tune(
                bool(adapt_step_size), d(acc.mean(dtype=d)),
                d(1.0 if init else 0.0))
        info = HMCInfo(samples=new_q, acceptance_rate=acc,
                       updated_step_size=self.step_size, init_momentum=p,
                       orig_hamiltonian=h0, hamiltonian=h1,
                       orig_log_prob=lp0, log_prob=new_lp,
                       step_size_used=eps, mass=mass, if_accept=if_accept)
        return new_q, info
