"""Importance-weighted objective (zhusuan/variational/monte_carlo.py:21-268):
``.tensor`` / ``.sgvb()`` on the K6 log_mean_exp kernel (forward + softmax
backward)."""
import torch

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
        """monte_carlo.py:166-227 (host-composed; a "next" row).
        O(K) per datum instead of the reference's [.., K, K] tile."""
        log_w = self._log_joint_term() + self._entropy_term()
        ax = self._axis
        K = log_w.shape[ax]
        if K < 2:
            raise ValueError(
                "VIMCO is a multi-sample gradient estimator, size along "
                "`axis` in the objective should be larger than 1.")
        l = log_w.detach()
        mean_except = (l.sum(ax, keepdim=True) - l) / (K - 1)
        # log_mean_exp with entry k replaced by mean_except[k]
        m = torch.maximum(l.max(ax, keepdim=True).values,
                          mean_except.max(ax, keepdim=True).values)
        s = torch.exp(l - m).sum(ax, keepdim=True)
        cv = torch.log((s - torch.exp(l - m) + torch.exp(mean_except - m))
                       / K) + m
        lme = ops.reduce_axes(log_w, ops.OP_LME, ax, keepdims=True)
        l_signal = (lme - cv).detach()
        fake_term = (-self._entropy_term() * l_signal).sum(ax)
        return -fake_term - lme.squeeze(ax)


def importance_weighted_objective(meta_bn, observed, latent=None, axis=None,
                                  variational=None):
    """monte_carlo.py:230-264."""
    return ImportanceWeightedObjective(
        meta_bn, observed, latent=latent, axis=axis, variational=variational)


iw_objective = importance_weighted_objective
