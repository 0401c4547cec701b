"""Log-joint descriptors that ``HMC.sample`` recognises and runs on fused
kernels.  Each is also a plain ``log_joint(observed_dict)`` callable built
from torch ops, so it works on the generic path and as its own cross-check.
"""
import math

import numpy as np
import torch

__all__ = ["GaussianLogJoint", "BNNRegressionLogJoint", "linear",
           "linear_bernoulli_log_prob", "LinearBernoulli"]


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



# ---------------------------------------------------------------------------
# K8: dense layers of a VAE / BNN log-joint on tcgen05 (gemm_logjoint_tc.cu)
# ---------------------------------------------------------------------------
def _tc_split(t2d):
    """fp32 [rows, K] -> (fp16 planes [2, rows, Kp], scale float[4]) for the
    tensor-core dense kernels (zsb_split16_pad_f32)."""
    from ._lib import lib, ptr, stream
    t2d = t2d.detach().to(torch.float32).contiguous()
    rows, K = int(t2d.shape[0]), int(t2d.shape[1])
    Kp = lib.load().zsb_linear_tc_kpad(K)
    planes = torch.empty((2, rows, Kp), dtype=torch.float16, device=t2d.device)
    scale = torch.zeros(4, dtype=torch.float32, device=t2d.device)
    lib.call("zsb_split16_pad_f32", ptr(t2d), rows, K, ptr(planes), ptr(scale),
             stream())
    return planes, scale


def _tc_split_t(t2d):
    """fp32 [R, C] -> (fp16 planes [2, C, Rp] of the transpose, scale)."""
    from ._lib import lib, ptr, stream
    t2d = t2d.detach().to(torch.float32).contiguous()
    R, C = int(t2d.shape[0]), int(t2d.shape[1])
    Rp = lib.load().zsb_linear_tc_kpad(R)
    planes = torch.empty((2, C, Rp), dtype=torch.float16, device=t2d.device)
    scale = torch.zeros(4, dtype=torch.float32, device=t2d.device)
    lib.call("zsb_split16_pad_t_f32", ptr(t2d), R, C, ptr(planes), ptr(scale),
             stream())
    return planes, scale


def _tc_grad_input(g, W):
    """dh [R, K] = g [R, J] @ W [J, K] on the tensor cores."""
    wtp, wts = _tc_split(W.detach().t())
    gp, gs = _tc_split(g)
    return _tc_linear(0, wtp, wts, gp, gs, None, None, None, int(g.shape[0]),
                      int(W.shape[1]), int(W.shape[0]))


def _tc_grad_weight(g, h2):
    """dW [J, K] = g^T [J, R] @ h [R, K]: contraction over the rows, split-K
    over the CTA pairs."""
    htp, hts = _tc_split_t(h2)
    gtp, gts = _tc_split_t(g)
    return _tc_linear(0, htp, hts, gtp, gts, None, None, None, int(g.shape[1]),
                      int(h2.shape[1]), int(h2.shape[0]), split_k=True)


def _tc_linear(epi, wp, ws, hp, hs, bias, x, gout, R, J, K, relu=False,
               split_k=False):
    from ._lib import lib, ptr, stream
    dev = hp.device
    part = None
    if epi == 0 and split_k:
        slices = lib.load().zsb_linear_tc_slices(R, J, K)
        if slices > 1:
            part = torch.empty(slices * R * J, dtype=torch.float32, device=dev)
    if epi == 1:
        out = torch.empty(R, dtype=torch.float32, device=dev)
        part = torch.empty(lib.load().zsb_linear_tc_nparts(J) * R,
                           dtype=torch.float32, device=dev)
    else:
        out = torch.empty((R, J), dtype=torch.float32, device=dev)
    lib.call("zsb_linear_tc_f32", epi, ptr(wp), ptr(ws), ptr(hp), ptr(hs),
             ptr(bias), ptr(x), int(x.shape[0]) if x is not None else 0,
             ptr(gout), ptr(out), ptr(part), R, J, K, int(bool(relu)), stream())
    return out


class _Linear(torch.autograd.Function):
    """y = relu?(h W^T + b): forward and both backward products on the
    tcgen05 kernel at fp32 accuracy (epi 0; the weight gradient uses the
    transposed operand planes and split-K)."""

    @staticmethod
    def forward(ctx, h, W, b, relu):
        lead = h.shape[:-1]
        h2 = h.reshape(-1, h.shape[-1])
        R, K, J = int(h2.shape[0]), int(h2.shape[1]), int(W.shape[0])
        wp, ws = _tc_split(W)
        hp, hs = _tc_split(h2)
        bias = b.detach().to(torch.float32).contiguous() if b is not None else None
        y = _tc_linear(0, wp, ws, hp, hs, bias, None, None, R, J, K, relu)
        ctx.save_for_backward(h2, W, y if relu else None)
        ctx.meta = (lead, relu, b is not None)
        return y.reshape(tuple(lead) + (J,))

    @staticmethod
    def backward(ctx, gy):
        h2, W, y = ctx.saved_tensors
        lead, relu, has_b = ctx.meta
        g = gy.reshape(-1, gy.shape[-1])
        if relu:
            g = g * (y > 0)
        g = g.to(torch.float32).contiguous()
        need = ctx.needs_input_grad
        dh = _tc_grad_input(g, W).reshape(tuple(lead) + (W.shape[1],)) \
            if need[0] else None
        dW = _tc_grad_weight(g, h2) if need[1] else None
        db = g.sum(0) if (has_b and need[2]) else None
        return dh, dW, db, None


def linear(h, W, b=None, relu=False):
    """``relu?(h @ W.T + b)`` (``tf.layers.dense``) on the tcgen05 kernel."""
    return _Linear.apply(h, W, b, bool(relu))


class _LinearBernoulliLogProb(torch.autograd.Function):
    """sum_j Bernoulli(logits = h W^T + b).log_prob(x)[..., j] without ever
    writing the logits: forward = GEMM with the Bernoulli row-sum epilogue
    (epi 1); backward = the same GEMM with the d/dlogits epilogue (epi 2),
    then the input / weight gradient products on the same kernel."""

    @staticmethod
    def forward(ctx, h, W, b, x):
        lead = h.shape[:-1]
        h2 = h.reshape(-1, h.shape[-1])
        R, K, J = int(h2.shape[0]), int(h2.shape[1]), int(W.shape[0])
        x2 = x.reshape(-1, J).to(torch.float32).contiguous()
        if R % int(x2.shape[0]) != 0:
            raise ValueError("rows of the observation (%d) must divide the rows "
                             "of the activations (%d)" % (x2.shape[0], R))
        wp, ws = _tc_split(W)
        hp, hs = _tc_split(h2)
        bias = b.detach().to(torch.float32).contiguous() if b is not None else None
        lp = _tc_linear(1, wp, ws, hp, hs, bias, x2, None, R, J, K)
        ctx.save_for_backward(h2, W, bias, x2, wp, ws, hp, hs)
        ctx.meta = (lead, R, J, K, b is not None)
        return lp.reshape(tuple(lead))

    @staticmethod
    def backward(ctx, glp):
        h2, W, bias, x2, wp, ws, hp, hs = ctx.saved_tensors
        lead, R, J, K, has_b = ctx.meta
        g = glp.reshape(-1).to(torch.float32).contiguous()
        dl = _tc_linear(2, wp, ws, hp, hs, bias, x2, g, R, J, K)
        need = ctx.needs_input_grad
        dh = _tc_grad_input(dl, W).reshape(tuple(lead) + (K,)) \
            if need[0] else None
        dW = _tc_grad_weight(dl, h2) if need[1] else None
        db = dl.sum(0) if (has_b and need[2]) else None
        return dh, dW, db, None


def linear_bernoulli_log_prob(h, W, b, x):
    """log p(x | logits = h @ W.T + b) summed over the last axis, fused.
    ``x`` ([n_x, J], 0/1) is broadcast over the leading rows of ``h``
    (row r of h uses x[r % n_x]: the [particles, batch] layout of
    iwae.py:23-32 flattened)."""
    return _LinearBernoulliLogProb.apply(h, W, b, x)


class LinearBernoulli(object):
    """Drop-in for ``Bernoulli(logits=dense(h), group_ndims=1)`` as a
    distribution plugin (duck-typed contract of bn.py:96-115): ``log_prob`` runs
    the fused GEMM + Bernoulli epilogue; the logits are only materialised when
    the node is *sampled* rather than observed."""

    def __init__(self, h, W, b=None, dtype=torch.int32, group_ndims=1):
        if group_ndims != 1:
            raise ValueError("LinearBernoulli sums over the feature axis: "
                             "group_ndims must be 1")
        self._h, self._W, self._b = h, W, b
        self.dtype = dtype
        self.param_dtype = torch.float32
        self.is_continuous = False
        self.is_reparameterized = False
        self.group_ndims = 1

    @property
    def logits(self):
        return linear(self._h, self._W, self._b)

    def get_batch_shape(self):
        return torch.Size(tuple(self._h.shape[:-1]) + (int(self._W.shape[0]),))

    def get_value_shape(self):
        return torch.Size([])

    batch_shape = property(lambda self: self.get_batch_shape())
    value_shape = property(lambda self: self.get_value_shape())

    def sample(self, n_samples=None):
        from .distributions import Bernoulli
        return Bernoulli(self.logits, dtype=self.dtype).sample(n_samples)

    def log_prob(self, given):
        J = int(self._W.shape[0])
        lead = tuple(self._h.shape[:-1])
        g = given.reshape(-1, J)
        n_x = int(g.shape[0])
        rows = 1
        for d in lead:
            rows *= int(d)
        if tuple(given.shape) != lead + (J,):
            # suffix-broadcast observation (e.g. x [N, J] against h [K, N, H])
            if rows % n_x != 0 or tuple(given.shape[:-1]) != lead[len(lead) - (given.dim() - 1):]:
                raise ValueError("given %s is not a suffix-broadcast of the "
                                 "batch shape %s" % (tuple(given.shape), lead + (J,)))
        return linear_bernoulli_log_prob(self._h, self._W, self._b, g)

    def prob(self, given):
        return torch.exp(self.log_prob(given))
