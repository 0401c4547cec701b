"""Distribution base class -- the contract of zhusuan/distributions/base.py:17-332
(``sample`` / ``log_prob`` / ``prob``, batch & value shapes, ``group_ndims``)."""
import torch

from ..utils import convert_to_tensor

__all__ = ["Distribution"]


class Distribution(object):
    # The reference's contract (base.py:290-304): `_log_prob` returns the UN-grouped log density
    # and `log_prob` sums the last `group_ndims` axes.  The built-in classes fuse that sum into
    # their kernels and say so here; a user-defined subclass (the plugin case) keeps the
    # reference behaviour and gets the sum from the base class.
    _group_sum_in_log_prob = False

    def __init__(self, dtype, param_dtype, is_continuous, is_reparameterized,
                 use_path_derivative=False, group_ndims=0, **kwargs):
        if 'group_event_ndims' in kwargs:
            raise ValueError(
                "The argument `group_event_ndims` has been deprecated "
                "Please use `group_ndims` instead.")
        self._dtype = dtype
        self._param_dtype = param_dtype
        self._is_continuous = is_continuous
        self._is_reparameterized = is_reparameterized
        self._use_path_derivative = use_path_derivative
        if isinstance(group_ndims, torch.Tensor):
            if group_ndims.dim() != 0:
                raise ValueError(
                    "group_ndims should be a scalar (0-D Tensor).")
            group_ndims = int(group_ndims.item())
        if not isinstance(group_ndims, int):
            raise TypeError("group_ndims must be an int")
        if group_ndims < 0:
            raise ValueError("group_ndims must be non-negative.")
        self._group_ndims = group_ndims
        # in-kernel Philox stream: (seed, draw counter); see ops.reparam_normal
        self._seed = kwargs.get("seed", None)
        self._draws = 0

    dtype = property(lambda self: self._dtype)
    param_dtype = property(lambda self: self._param_dtype)
    is_continuous = property(lambda self: self._is_continuous)
    is_reparameterized = property(lambda self: self._is_reparameterized)
    use_path_derivative = property(lambda self: self._use_path_derivative)
    group_ndims = property(lambda self: self._group_ndims)

    def path_param(self, param):
        """base.py:150-157: detach params when using the path derivative."""
        return param.detach() if self._use_path_derivative else param

    # shapes are static in the torch world: both flavours return torch.Size
    @property
    def value_shape(self):
        return self._get_value_shape()

    def get_value_shape(self):
        return self._get_value_shape()

    @property
    def batch_shape(self):
        return self._get_batch_shape()

    def get_batch_shape(self):
        return self._get_batch_shape()

    def _next_rng(self):
        from .. import random as zrandom
        seed = self._seed if self._seed is not None else zrandom.get_seed()
        self._draws += 1
        return seed, zrandom.next_counter()

    def sample(self, n_samples=None):
        """base.py:236-263: None -> one sample with the leading axis squeezed."""
        if n_samples is None:
            return self._sample(1).squeeze(0)
        if isinstance(n_samples, torch.Tensor):
            if n_samples.dim() != 0:
                raise ValueError("n_samples should be a scalar (0-D Tensor).")
            n_samples = int(n_samples.item())
        if not isinstance(n_samples, int):
            raise TypeError("n_samples must be an int or a 0-D integer tensor")
        return self._sample(n_samples)

    def _check_input_shape(self, given):
        """base.py:271-288: static broadcast check against batch + value shape."""
        given = convert_to_tensor(given, dtype=self.dtype)
        err_msg = "The given argument should be able to broadcast to " \
                  "match batch_shape + value_shape of the distribution."
        sample_shape = tuple(self.get_batch_shape()) + tuple(
            self.get_value_shape())
        try:
            torch.broadcast_shapes(tuple(given.shape), sample_shape)
        except RuntimeError:
            raise ValueError(err_msg + " ({} vs. {} + {})".format(
                tuple(given.shape), tuple(self.get_batch_shape()),
                tuple(self.get_value_shape())))
        return given

    def log_prob(self, given):
        """base.py:290-304: shape check, `_log_prob`, sum over the last
        `group_ndims` axes (fused into the kernels of the built-in classes)."""
        given = self._check_input_shape(given)
        lp = self._log_prob(given)
        if not self._group_sum_in_log_prob and self._group_ndims > 0:
            from .. import ops
            lp = ops.group_sum(lp, self._group_ndims)
        return lp

    def prob(self, given):
        """base.py:306-320 (exp of the grouped log-prob == prod of probs)."""
        return torch.exp(self.log_prob(given))

    def _get_value_shape(self):
        raise NotImplementedError()

    def _get_batch_shape(self):
        raise NotImplementedError()

    def _sample(self, n_samples):
        raise NotImplementedError()

    def _log_prob(self, given):
        raise NotImplementedError()
