"""The other elementwise univariate distributions of
zhusuan/distributions/univariate.py (FoldNormal 187-331, Uniform 557-659,
Gamma 662-750, Beta 753-854, Poisson 857-936, Binomial 939-1067,
InverseGamma 1070-1161, Laplace 1164-1276, BinConcrete 1279-1405).

``log_prob`` (+ gradients wrt the value and both parameters) runs in the
zsb_logprob_univariate_f32 kernels.  Sampling is off the accelerated path: it
uses torch's device-side samplers (uniform / gamma / poisson / binomial) with
the same transformations as the reference's ``_sample``.
"""
import numpy as np
import torch

from .. import ops
from ..utils import convert_to_tensor
from .base import Distribution
from .utils import (assert_same_float_dtype, assert_dtype_is_int_or_float,
                    broadcast_check)

__all__ = ["FoldNormal", "Uniform", "Gamma", "Beta", "Poisson", "Binomial",
           "InverseGamma", "Laplace", "BinConcrete", "BinGumbelSoftmax"]


def _pair(cls_name, a, a_name, b, b_name):
    """Convert two parameters, check dtype agreement and broadcastability."""
    a = convert_to_tensor(a)
    b = convert_to_tensor(b, device=a.device)
    dtype = assert_same_float_dtype([(a, '%s.%s' % (cls_name, a_name)),
                                     (b, '%s.%s' % (cls_name, b_name))])
    broadcast_check(a.shape, b.shape,
                    "{} and {} should be broadcastable to match each "
                    "other. ({} vs. {})".format(a_name, b_name, tuple(a.shape),
                                                tuple(b.shape)))
    return a, b, dtype


def _checked(lp, flag, what):
    if flag and not bool(torch.isfinite(lp).all()):
        raise FloatingPointError("%s has numeric errors" % what)
    return lp


def _sample_shape(n_samples, batch_shape):
    return (int(n_samples),) + tuple(batch_shape)


class _TwoParam(Distribution):
    """Scalar-valued distribution whose batch shape is the broadcast of two
    parameters held in ``self._a`` / ``self._b``."""
    _group_sum_in_log_prob = True
    _dist_id = None

    def _get_value_shape(self):
        return torch.Size([])

    def _get_batch_shape(self):
        return torch.broadcast_shapes(self._a.shape, self._b.shape)

    def _params(self):
        return self._a, self._b

    def _log_prob(self, given):
        a, b = self._params()
        lp = ops.univariate_log_prob(self._dist_id, given, a, b,
                                     self._group_ndims)
        return _checked(lp, self._check_numerics, type(self).__name__ + ".log_prob")


class FoldNormal(_TwoParam):
    """univariate.py:187-331: |N(mean, std)|; ``std`` xor ``logstd``."""
    _dist_id = ops.UNI_FOLDNORMAL

    def __init__(self, mean=0., _sentinel=None, std=None, logstd=None,
                 group_ndims=0, is_reparameterized=True,
                 use_path_derivative=False, check_numerics=False, **kwargs):
        if _sentinel is not None:
            raise ValueError(
                "The order of logstd/std has changed to std/logstd since "
                "0.3.1. Please use named arguments: FoldNormal(mean, std=..., "
                "...) or FoldNormal(mean, logstd=..., ...).")
        if (logstd is None) == (std is None):
            raise ValueError("Either std or logstd should be passed but not "
                             "both of them.")
        if logstd is None:
            self._mean, self._std, dtype = _pair("FoldNormal", mean, "mean",
                                                 std, "std")
            self._logstd = torch.log(self._std)
        else:
            self._mean, self._logstd, dtype = _pair("FoldNormal", mean, "mean",
                                                    logstd, "logstd")
            self._std = torch.exp(self._logstd)
        self._a, self._b = self._mean, self._logstd
        self._check_numerics = check_numerics
        super(FoldNormal, self).__init__(
            dtype=dtype, param_dtype=dtype, is_continuous=True,
            is_reparameterized=is_reparameterized,
            use_path_derivative=use_path_derivative, group_ndims=group_ndims,
            **kwargs)

    mean = property(lambda self: self._mean)
    logstd = property(lambda self: self._logstd)
    std = property(lambda self: self._std)

    def _params(self):
        return self.path_param(self._mean), self.path_param(self._logstd)

    def _sample(self, n_samples):
        # univariate.py:306-317 (the reference returns the un-folded normal draw)
        mean, std = self._mean, self._std
        if not self.is_reparameterized:
            mean, std = mean.detach(), std.detach()
        eps = ops.base_noise(1, _sample_shape(n_samples, self._get_batch_shape()), mean.device,
                             *self._next_rng())
        return eps * std + mean


class Uniform(_TwoParam):
    """univariate.py:557-659 on [minval, maxval)."""
    _dist_id = ops.UNI_UNIFORM

    def __init__(self, minval=0., maxval=1., group_ndims=0,
                 is_reparameterized=True, check_numerics=False, **kwargs):
        self._a, self._b, dtype = _pair("Uniform", minval, "minval", maxval,
                                        "maxval")
        self._check_numerics = check_numerics
        super(Uniform, self).__init__(
            dtype=dtype, param_dtype=dtype, is_continuous=True,
            is_reparameterized=is_reparameterized, group_ndims=group_ndims,
            **kwargs)

    minval = property(lambda self: self._a)
    maxval = property(lambda self: self._b)

    def _sample(self, n_samples):
        lo, hi = self._a, self._b
        if not self.is_reparameterized:
            lo, hi = lo.detach(), hi.detach()
        u = ops.base_noise(0, _sample_shape(n_samples, self._get_batch_shape()), lo.device,
                           *self._next_rng())
        return u * (hi - lo) + lo


class Gamma(_TwoParam):
    """univariate.py:662-750: shape alpha, rate beta."""
    _dist_id = ops.UNI_GAMMA

    def __init__(self, alpha, beta, group_ndims=0, check_numerics=False,
                 **kwargs):
        self._a, self._b, dtype = _pair(type(self).__name__, alpha, "alpha",
                                        beta, "beta")
        self._check_numerics = check_numerics
        super(Gamma, self).__init__(
            dtype=dtype, param_dtype=dtype, is_continuous=True,
            is_reparameterized=False, group_ndims=group_ndims, **kwargs)

    alpha = property(lambda self: self._a)
    beta = property(lambda self: self._b)

    def _gamma_draw(self, n_samples, conc):
        # Gamma(conc, 1) on the device sampler (Marsaglia-Tsang on Philox, csrc/samplers.cu)
        shape = _sample_shape(n_samples, self._get_batch_shape())
        seed, it = self._next_rng()
        return ops.sample_gamma(conc, None, shape, seed=seed, it=it)

    def _sample(self, n_samples):
        return self._gamma_draw(n_samples, self._a) / self._b.detach()


class InverseGamma(Gamma):
    """univariate.py:1070-1161: 1 / Gamma(alpha, beta)."""
    _dist_id = ops.UNI_INVGAMMA

    def _sample(self, n_samples):
        return self._b.detach() / self._gamma_draw(n_samples, self._a)


class Beta(Gamma):
    """univariate.py:753-854."""
    _dist_id = ops.UNI_BETA

    def _sample(self, n_samples):
        x = self._gamma_draw(n_samples, self._a)
        y = self._gamma_draw(n_samples, self._b)
        return x / (x + y)


class Laplace(_TwoParam):
    """univariate.py:1164-1276."""
    _dist_id = ops.UNI_LAPLACE

    def __init__(self, loc, scale, group_ndims=0, is_reparameterized=True,
                 use_path_derivative=False, check_numerics=False, **kwargs):
        self._a, self._b, dtype = _pair("Laplace", loc, "loc", scale, "scale")
        self._check_numerics = check_numerics
        super(Laplace, self).__init__(
            dtype=dtype, param_dtype=dtype, is_continuous=True,
            is_reparameterized=is_reparameterized,
            use_path_derivative=use_path_derivative, group_ndims=group_ndims,
            **kwargs)

    loc = property(lambda self: self._a)
    scale = property(lambda self: self._b)

    def _params(self):
        return self.path_param(self._a), self.path_param(self._b)

    def _sample(self, n_samples):
        loc, scale = self._a, self._b
        if not self.is_reparameterized:
            loc, scale = loc.detach(), scale.detach()
        # u in (-1, 1): inverse-CDF draw (univariate.py:1246-1265)
        u = ops.base_noise(0, _sample_shape(n_samples, self._get_batch_shape()), loc.device,
                           *self._next_rng())
        u = (2.0 * u - 1.0).clamp(min=-1.0 + 2.0 ** -24)
        return loc - scale * torch.sign(u) * torch.log1p(-torch.abs(u))


class BinConcrete(_TwoParam):
    """univariate.py:1279-1405 (binary Gumbel-softmax relaxation)."""
    _dist_id = ops.UNI_BINCONCRETE

    def __init__(self, temperature, logits, group_ndims=0,
                 is_reparameterized=True, use_path_derivative=False,
                 check_numerics=False, **kwargs):
        self._logits = convert_to_tensor(logits)
        self._temperature = convert_to_tensor(temperature,
                                              device=self._logits.device)
        dtype = assert_same_float_dtype(
            [(self._logits, 'BinConcrete.logits'),
             (self._temperature, 'BinConcrete.temperature')])
        if self._temperature.dim() != 0:
            raise ValueError("BinConcrete.temperature should be a scalar "
                             "(0-D Tensor).")
        self._a, self._b = self._temperature, self._logits
        self._check_numerics = check_numerics
        super(BinConcrete, self).__init__(
            dtype=dtype, param_dtype=dtype, is_continuous=True,
            is_reparameterized=is_reparameterized,
            use_path_derivative=use_path_derivative, group_ndims=group_ndims,
            **kwargs)

    temperature = property(lambda self: self._temperature)
    logits = property(lambda self: self._logits)

    def _get_batch_shape(self):
        return self._logits.shape

    def _params(self):
        return self.path_param(self._temperature), self.path_param(self._logits)

    def _sample(self, n_samples):
        logits, temperature = self._logits, self._temperature
        if not self.is_reparameterized:
            logits, temperature = logits.detach(), temperature.detach()
        u = ops.base_noise(0, _sample_shape(n_samples, logits.shape), logits.device,
                           *self._next_rng()).clamp(1e-7, 1.0 - 1e-7)
        logistic = torch.log(u) - torch.log1p(-u)
        return torch.sigmoid((logits + logistic) / temperature)


BinGumbelSoftmax = BinConcrete


class Poisson(Distribution):
    """univariate.py:857-936."""
    _group_sum_in_log_prob = True

    def __init__(self, rate, dtype=torch.int32, group_ndims=0,
                 check_numerics=False, **kwargs):
        self._rate = convert_to_tensor(rate)
        param_dtype = assert_same_float_dtype([(self._rate, 'Poisson.rate')])
        assert_dtype_is_int_or_float(dtype)
        self._check_numerics = check_numerics
        super(Poisson, self).__init__(
            dtype=dtype, param_dtype=param_dtype, is_continuous=False,
            is_reparameterized=False, group_ndims=group_ndims, **kwargs)

    rate = property(lambda self: self._rate)

    def _get_value_shape(self):
        return torch.Size([])

    def _get_batch_shape(self):
        return self._rate.shape

    def _sample(self, n_samples, u=None):
        # tf.random_poisson (univariate.py:915-920) -> device sampler (inverse transform from the
        # mode, one Philox uniform per draw; ``u``: injected uniforms for parity)
        seed, it = self._next_rng()
        out = ops.sample_count(0, self._rate, 0, _sample_shape(n_samples, self._rate.shape),
                               u=u, seed=seed, it=it)
        return out if self.dtype == torch.int32 else out.to(self.dtype)

    def _log_prob(self, given):
        lp = ops.univariate_log_prob(ops.UNI_POISSON, given, self._rate, None,
                                     self._group_ndims)
        return _checked(lp, self._check_numerics, "Poisson.log_prob")


class Binomial(Distribution):
    """univariate.py:939-1067."""
    _group_sum_in_log_prob = True

    def __init__(self, logits, n_experiments, dtype=torch.int32, group_ndims=0,
                 check_numerics=False, **kwargs):
        self._logits = convert_to_tensor(logits)
        param_dtype = assert_same_float_dtype(
            [(self._logits, 'Binomial.logits')])
        assert_dtype_is_int_or_float(dtype)
        if isinstance(n_experiments, torch.Tensor):
            if n_experiments.dtype not in (torch.int32, torch.int64):
                raise TypeError('n_experiments must be int32')
            if n_experiments.dim() != 0:
                raise ValueError(
                    "n_experiments should be a scalar (0-D Tensor).")
            n_experiments = int(n_experiments)
        elif not isinstance(n_experiments, (int, np.integer)):
            raise TypeError('n_experiments must be int32')
        if n_experiments <= 0:
            raise ValueError("n_experiments must be positive")
        self._n_experiments = int(n_experiments)
        self._n_f = torch.tensor(float(self._n_experiments),
                                 dtype=torch.float32,
                                 device=self._logits.device)
        self._check_numerics = check_numerics
        super(Binomial, self).__init__(
            dtype=dtype, param_dtype=param_dtype, is_continuous=False,
            is_reparameterized=False, group_ndims=group_ndims, **kwargs)

    n_experiments = property(lambda self: self._n_experiments)
    logits = property(lambda self: self._logits)

    def _get_value_shape(self):
        return torch.Size([])

    def _get_batch_shape(self):
        return self._logits.shape

    def _sample(self, n_samples, u=None):
        # univariate.py:1025-1045 (n categorical draws summed) -> device sampler: the count itself
        # by inverse transform from the mode, one Philox uniform per draw
        seed, it = self._next_rng()
        out = ops.sample_count(1, self._logits, self._n_experiments,
                               _sample_shape(n_samples, self._logits.shape), u=u, seed=seed, it=it)
        return out if self.dtype == torch.int32 else out.to(self.dtype)

    def _log_prob(self, given):
        lp = ops.univariate_log_prob(ops.UNI_BINOMIAL, given, self._logits,
                                     self._n_f, self._group_ndims)
        return _checked(lp, self._check_numerics, "Binomial.log_prob")
