"""Log-joint descriptors that ``HMC.sample`` recognises and runs on fused
kernels.  Each is also a plain ``log_joint(observed_dict)`` callable built
from torch ops, so it works on the generic path and as its own cross-check.
"""
import math

import numpy as np
import torch

__all__ = ["GaussianLogJoint", "BNNRegressionLogJoint"]


class GaussianLogJoint(object):
    """log p(x) = -1/2 (x-mu)^T P (x-mu) - 1/2 log|2 pi Sigma|, P = Sigma^-1
    shared by all chains (BASELINE config 2; the reference can only express
    this as a callable ``log_joint`` because MultivariateNormalCholesky
    broadcasts L to every chain, multivariate.py:183-185).

    precision: [D, D] symmetric (numpy float64 preferred: the b = P mu vector
    and the hi/lo split are derived in float64 on the host, once).
    """

    def __init__(self, precision, mean=None, log_det_cov=None, name="x",
                 device="cuda", impl=None):
        P64 = np.asarray(precision.detach().cpu().numpy()
                         if isinstance(precision, torch.Tensor)
                         else precision, dtype=np.float64)
        D = P64.shape[0]
        if P64.shape != (D, D):
            raise ValueError("precision must be square")
        if D % 16 != 0:
            raise ValueError("the fused dense-Gaussian path needs D % 16 == 0")
        P64 = 0.5 * (P64 + P64.T)
        self.name = name
        self.D = D
        if log_det_cov is None:
            sign, ld = np.linalg.slogdet(P64)
            log_det_cov = -ld
        self.const = float(-0.5 * (D * math.log(2 * math.pi) + log_det_cov))
        P32 = P64.astype(np.float32)
        self.P = torch.as_tensor(P32, device=device).contiguous()
        d = {"kind": "dense_gaussian", "D": D, "P": self.P,
             "const": self.const, "impl": impl}
        self.mu = None
        if mean is not None:
            mu64 = np.asarray(mean, np.float64).reshape(D)
            self.mu = torch.as_tensor(mu64.astype(np.float32), device=device)
            d["mu"] = self.mu
            d["b"] = torch.as_tensor(
                (P32.astype(np.float64) @ mu64).astype(np.float32),
                device=device)
        # 3xTF32 split for the tensor-core path: hi = fp32 with the low 13
        # mantissa bits cleared (exactly representable in TF32), lo = P - hi.
        hi = (P32.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        lo = (P32 - hi).astype(np.float32)
        d["P_hi"] = torch.as_tensor(hi, device=device).contiguous()
        d["P_lo"] = torch.as_tensor(lo, device=device).contiguous()
        # fp16 split for impl 2: P*sP = h + l with sP a power of two that puts
        # max|P| in [2^11, 2^12); products then fit fp32 accumulation exactly.
        pmax = float(np.abs(P32).max())
        sP = float(2.0 ** (12 - np.frexp(pmax)[1])) if pmax > 0 else 1.0
        x = P32.astype(np.float64) * sP
        h16 = x.astype(np.float16)
        l16 = (x - h16.astype(np.float64)).astype(np.float16)
        d["sP"] = sP
        d["P_h16"] = torch.as_tensor(h16, device=device).contiguous()
        d["P_l16"] = torch.as_tensor(l16, device=device).contiguous()
        self._zsb_fused = d

    def __call__(self, observed):
        x = observed[self.name]
        if self.mu is not None:
            x = x - self.mu
        return -0.5 * ((x @ self.P) * x).sum(-1) + self.const


class BNNRegressionLogJoint(object):
    """The log-joint of examples/bayesian_neural_nets/bnn_sgmcmc.py:19-35, 74-77
    for layer sizes [n_in, H, 1] and per-chain weights:

        w0 [chains, H, n_in+1] ~ N(0, exp(logstds[0])),
        w1 [chains, 1, H+1]    ~ N(0, exp(logstds[1])),
        y ~ N(net(x; w), exp(y_logstd)),
        log_joint = sum log p(w) + mean_batch(log p(y|x,w)) * n_train.

    As a callable it is the generic-path log-joint (registry Normal kernels +
    torch einsum under the tape); ``zs.SGHMC.sample`` recognises it and runs
    the whole step in one fused kernel (zsb_sgmcmc_sghmc_bnn_f32).  Feed
    minibatches with ``sample_op(observed={'x': xb, 'y': yb})``.
    """

    def __init__(self, x, y, logstds, n_train, y_logstd=-0.95,
                 names=("w0", "w1")):
        from .distributions import Normal
        self._Normal = Normal
        self.x, self.y = x, y
        self.logstds = [l.contiguous() for l in logstds]
        self.n_train = float(n_train)
        self.y_logstd = float(y_logstd)
        self.names = tuple(names)
        self._zsb_fused = {"kind": "bnn_regression", "obj": self}

    def set_batch(self, observed):
        if "x" in observed:
            self.x = observed["x"]
        if "y" in observed:
            self.y = observed["y"]

    def __call__(self, observed):
        x = observed.get("x", self.x)
        y = observed.get("y", self.y)
        w0, w1 = observed[self.names[0]], observed[self.names[1]]
        C = w0.shape[0]
        h = x.unsqueeze(0).expand(C, -1, -1)
        lp = 0.0
        for w, ls in zip((w0, w1), self.logstds):
            ones = torch.ones(h.shape[:-1] + (1,), device=h.device)
            h = torch.cat([h, ones], -1)
            h = torch.einsum("imk,ijk->ijm", w, h) / math.sqrt(h.shape[2])
            if w is w0:
                h = torch.relu(h)
            lp = lp + self._Normal(torch.zeros_like(ls), logstd=ls,
                                   group_ndims=2).log_prob(w)
        y_mean = h.squeeze(2)
        lpy = self._Normal(y_mean, logstd=torch.full_like(y_mean,
                                                         self.y_logstd)
                           ).log_prob(y.unsqueeze(0))
        return lp + lpy.mean(1) * self.n_train
