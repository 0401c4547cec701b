"""NumPy restatement of zhusuan/sgmcmc.py (TEST ORACLE ONLY).

All ``tf.random_normal`` draws are INJECTED as standard-normal arrays and
scaled by the reference's stddev inside the update, in the reference's order
(see each method).  ``t`` is the int32 iteration counter (sgmcmc.py:73); the
resample test ``t % n_iter_resample_v == 0`` is evaluated on the
pre-increment value (first call: t == 0), which is the order the oracle
fixes for the reference's otherwise unspecified read/increment race
(sgmcmc.py:107-108 vs 335).
"""
import numpy as np


class _Base(object):
    def __init__(self, dtype=np.float32):
        self.dtype = dtype
        self.t = 0                                     # sgmcmc.py:73


class SGLD(_Base):
    """sgmcmc.py:170-200:  q += 0.5*lr*g + N(0, sqrt(lr))."""

    def __init__(self, learning_rate, dtype=np.float32):
        super(SGLD, self).__init__(dtype)
        self.lr = dtype(learning_rate)

    def step(self, qs, grad, noise):
        d = self.dtype
        gs = grad(qs)
        out = []
        for q, g, n in zip(qs, gs, noise):
            out.append((q + d(0.5) * self.lr * g.astype(d)
                        + np.asarray(n, d) * np.sqrt(self.lr)).astype(d))
        self.t += 1
        return out, {"q": out}


class PSGLD(SGLD):
    """sgmcmc.py:203-257 (RMSprop preconditioner, decay 0.9, eps 1e-3)."""

    def __init__(self, learning_rate, decay=0.9, epsilon=1e-3,
                 dtype=np.float32):
        super(PSGLD, self).__init__(learning_rate, dtype)
        self.decay, self.epsilon = dtype(decay), dtype(epsilon)
        self.aux = None

    def step(self, qs, grad, noise):
        d = self.dtype
        if self.aux is None:
            self.aux = [np.zeros_like(q, d) for q in qs]    # :225-226
        gs = grad(qs)
        out = []
        for k, (q, g, n) in enumerate(zip(qs, gs, noise)):
            g = g.astype(d)
            self.aux[k] = (self.decay * self.aux[k]
                           + (d(1) - self.decay) * g * g).astype(d)  # :230
            G = (d(1) / (self.epsilon + np.sqrt(self.aux[k]))).astype(d)
            out.append((q + d(0.5) * self.lr * G * g
                        + np.asarray(n, d) * np.sqrt(self.lr * G)).astype(d))
        self.t += 1
        return out, {"q": out}


class SGHMC(_Base):
    """sgmcmc.py:260-371."""

    def __init__(self, learning_rate, friction=0.25, variance_estimate=0.,
                 n_iter_resample_v=20, second_order=True, dtype=np.float32):
        super(SGHMC, self).__init__(dtype)
        d = dtype
        self.lr, self.alpha, self.beta = (d(learning_rate), d(friction),
                                          d(variance_estimate))
        self.n_iter_resample_v = int(n_iter_resample_v or 0)
        self.second_order = second_order
        self.vs = None

    def init_v(self, noise_v0):
        """sgmcmc.py:320-324: v0 ~ N(0, sqrt(lr))."""
        d = self.dtype
        self.vs = [(np.asarray(n, d) * np.sqrt(self.lr)).astype(d)
                   for n in noise_v0]

    def step(self, qs, grad, noise_resample, noise):
        """noise_resample: std-normals used iff the resample branch fires;
        noise: std-normals for the injected gaussian term."""
        d = self.dtype
        resample = (self.n_iter_resample_v != 0
                    and self.t % self.n_iter_resample_v == 0)    # :330-336
        old_vs = [(np.asarray(n, d) * np.sqrt(self.lr)).astype(d)
                  if resample else v
                  for v, n in zip(self.vs, noise_resample)]
        std = np.sqrt(d(2) * (self.alpha - self.beta) * self.lr)  # :341
        terms = [(np.asarray(n, d) * std).astype(d) for n in noise]
        if not self.second_order:                                # :343-348
            gs = grad(qs)
            new_vs = [((d(1) - self.alpha) * v + self.lr * g.astype(d)
                       + t).astype(d) for v, g, t in zip(old_vs, gs, terms)]
            new_qs = [(q + v).astype(d) for q, v in zip(qs, new_vs)]
        else:                                                    # :349-356
            dh = d(np.exp(d(-0.5) * self.alpha))
            q1s = [(q + d(0.5) * v).astype(d) for q, v in zip(qs, old_vs)]
            gs = grad(q1s)
            new_vs = [(dh * (dh * v + self.lr * g.astype(d) + t)).astype(d)
                      for v, g, t in zip(old_vs, gs, terms)]
            new_qs = [(q1 + d(0.5) * v).astype(d)
                      for q1, v in zip(q1s, new_vs)]
        mean_ks = [d(np.mean(v * v, dtype=d)) for v in new_vs]   # :358
        self.vs = new_vs
        self.t += 1
        return new_qs, {"q": new_qs, "mean_k": mean_ks}


class SGNHT(_Base):
    """sgmcmc.py:374-523."""

    def __init__(self, learning_rate, variance_extra=0., tune_rate=1.,
                 n_iter_resample_v=None, second_order=True,
                 use_vector_alpha=True, dtype=np.float32):
        super(SGNHT, self).__init__(dtype)
        d = dtype
        self.lr, self.a, self.tune_rate = (d(learning_rate),
                                           d(variance_extra), d(tune_rate))
        self.n_iter_resample_v = int(n_iter_resample_v or 0)
        self.second_order = second_order
        self.use_vector_alpha = use_vector_alpha
        self.vs = None
        self.alphas = None

    def init_v(self, noise_v0):
        d = self.dtype
        self.vs = [(np.asarray(n, d) * np.sqrt(self.lr)).astype(d)
                   for n in noise_v0]                            # :450-452
        if self.use_vector_alpha:                                # :454-458
            self.alphas = [(self.a * np.ones_like(v)).astype(d)
                           for v in self.vs]
        else:
            self.alphas = [d(self.a) for _ in self.vs]

    def _mrm(self, x):
        return x if self.use_vector_alpha else self.dtype(
            np.mean(x, dtype=self.dtype))

    def step(self, qs, grad, noise_resample, noise):
        d = self.dtype
        resample = (self.n_iter_resample_v != 0
                    and self.t % self.n_iter_resample_v == 0)
        old_vs = [(np.asarray(n, d) * np.sqrt(self.lr)).astype(d)
                  if resample else v
                  for v, n in zip(self.vs, noise_resample)]
        std = np.sqrt(d(2) * self.a * self.lr)                   # :480
        terms = [(np.asarray(n, d) * std).astype(d) for n in noise]
        if not self.second_order:                                # :483-491
            gs = grad(qs)
            new_vs = [((d(1) - al) * v + self.lr * g.astype(d) + t).astype(d)
                      for v, al, g, t in zip(old_vs, self.alphas, gs, terms)]
            new_qs = [(q + v).astype(d) for q, v in zip(qs, new_vs)]
            mean_ks = [self._mrm(v * v) for v in new_vs]
            new_alphas = [al + self.tune_rate * (mk - self.lr)
                          for al, mk in zip(self.alphas, mean_ks)]
        else:                                                    # :492-507
            q1s = [(q + d(0.5) * v).astype(d) for q, v in zip(qs, old_vs)]
            mk1 = [self._mrm(v * v) for v in old_vs]
            a1s = [al + d(0.5) * self.tune_rate * (m - self.lr)
                   for al, m in zip(self.alphas, mk1)]
            dhs = [np.exp(d(-0.5) * a1).astype(d) for a1 in a1s]
            gs = grad(q1s)
            new_vs = [(dh * (dh * v + self.lr * g.astype(d) + t)).astype(d)
                      for dh, v, g, t in zip(dhs, old_vs, gs, terms)]
            new_qs = [(q1 + d(0.5) * v).astype(d)
                      for q1, v in zip(q1s, new_vs)]
            mean_ks = [self._mrm(v * v) for v in new_vs]
            new_alphas = [a1 + d(0.5) * self.tune_rate * (mk - self.lr)
                          for a1, mk in zip(a1s, mean_ks)]
        self.vs = new_vs
        self.alphas = [np.asarray(a, d) if self.use_vector_alpha else d(a)
                       for a in new_alphas]
        self.t += 1
        return new_qs, {"q": new_qs, "mean_k": mean_ks,
                        "alpha": self.alphas}
