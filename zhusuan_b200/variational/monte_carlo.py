"""Importance-weighted objective (zhusuan/variational/monte_carlo.py:21-268):
``.tensor`` / ``.sgvb()`` on the K6 log_mean_exp kernel (forward + softmax
backward)."""
from .. import ops
from .base import VariationalObjective

__all__ = ['importance_weighted_objective', 'iw_objective',
           'ImportanceWeightedObjective']


class ImportanceWeightedObjective(VariationalObjective):
    def __init__(self, meta_bn, observed, latent=None, axis=None,
                 variational=None):
        if axis is None:                           # monte_carlo.py:126-129
            raise ValueError(
                "ImportanceWeightedObjective is a multi-sample objective, "
                "the `axis` argument must be specified.")
        self._axis = axis
        super(ImportanceWeightedObjective, self).__init__(
            meta_bn, observed, latent=latent, variational=variational)

    def _objective(self):                          # monte_carlo.py:137-141
        log_w = self._log_joint_term() + self._entropy_term()
        if self._axis is not None:
            return ops.reduce_axes(log_w, ops.OP_LME, self._axis)
        return log_w

    def sgvb(self):                                # monte_carlo.py:143-164
        return -self.tensor

    def vimco(self):
        """monte_carlo.py:166-227.  The learning signal
        LME(log_w) - LME(log_w with entry k replaced by the mean of the
        others) comes from one O(K)-per-datum kernel (zsb_vimco_signal_f32)
        instead of the reference's [.., K, K] tile; gradients flow through
        log q (the fake term) and through log_mean_exp(log_w) as in the
        reference."""
        log_w = self._log_joint_term() + self._entropy_term()
        l_signal, _ = ops.vimco_signal(log_w, self._axis)   # ValueError if K < 2
        fake_term = ops.reduce_axes(-self._entropy_term() * l_signal,
                                    ops.OP_SUM, self._axis)
        return -fake_term - ops.reduce_axes(log_w, ops.OP_LME, self._axis)


def importance_weighted_objective(meta_bn, observed, latent=None, axis=None,
                                  variational=None):
    """monte_carlo.py:230-264."""
    return ImportanceWeightedObjective(
        meta_bn, observed, latent=latent, axis=axis, variational=variational)


iw_objective = importance_weighted_objective
