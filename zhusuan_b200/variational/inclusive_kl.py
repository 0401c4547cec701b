"""Inclusive KL objective KL(p || q) (zhusuan/variational/inclusive_kl.py:20-186):
optimised through the self-normalised importance-sampling estimator (``importance()``,
formerly ``rws()``); the normalised weights come from zsb_normalized_weights_f32."""
import warnings

from .. import ops
from .base import VariationalObjective

__all__ = ['klpq', 'InclusiveKLObjective']


class InclusiveKLObjective(VariationalObjective):
    def __init__(self, meta_bn, observed, latent=None, axis=None,
                 variational=None):
        self._axis = axis
        super(InclusiveKLObjective, self).__init__(
            meta_bn, observed, latent=latent, variational=variational)

    def _objective(self):                          # inclusive_kl.py:104-107
        raise NotImplementedError(
            "The inclusive KL objective (klpq) can only be optimized instead "
            "of being evaluated.")

    def rws(self):                                 # inclusive_kl.py:109-117
        warnings.warn(
            "The `rws()` method has been renamed to `importance()`, "
            "`rws()` will be removed in the coming version (0.4.1)",
            FutureWarning)
        return self.importance()

    def importance(self):
        """inclusive_kl.py:119-151: cost = sum_axis(w~ * (-log q)), w~ the
        (constant) self-normalised importance weights."""
        entropy = self._entropy_term()
        if self._axis is None:
            warnings.warn(
                "The gradient estimator is using self-normalized "
                "importance sampling, which is heavily biased and inaccurate "
                "when you're using only a single sample (`axis=None`).")
            return entropy
        log_w = self._log_joint_term() + entropy
        w_tilde = ops.normalized_weights(log_w, self._axis)
        return ops.reduce_axes(w_tilde * entropy, ops.OP_SUM, self._axis)


def klpq(meta_bn, observed, latent=None, axis=None, variational=None):
    """inclusive_kl.py:154-186."""
    return InclusiveKLObjective(
        meta_bn, observed, latent=latent, axis=axis, variational=variational)
