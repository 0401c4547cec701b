"""ELBO objective (zhusuan/variational/exclusive_kl.py:20-267): ``.tensor``
and ``.sgvb()`` on the K6 ``mean`` reduction kernel (fwd + 1/K backward)."""
import torch

from .. import ops
from .base import VariationalObjective

__all__ = ['elbo', 'EvidenceLowerBoundObjective']


class EvidenceLowerBoundObjective(VariationalObjective):
    def __init__(self, meta_bn, observed, latent=None, axis=None,
                 variational=None):
        self._axis = axis
        super(EvidenceLowerBoundObjective, self).__init__(
            meta_bn, observed, latent=latent, variational=variational)

    def _objective(self):                          # exclusive_kl.py:131-137
        lower_bound = self._log_joint_term()
        if self._entropy_term() is not None:
            lower_bound = lower_bound + self._entropy_term()
        if self._axis is not None:
            lower_bound = ops.reduce_axes(lower_bound, ops.OP_MEAN,
                                          self._axis)
        return lower_bound

    def sgvb(self):                                # exclusive_kl.py:139-159
        return -self.tensor

    def reinforce(self, variance_reduction=True, baseline=None, decay=0.8):
        """exclusive_kl.py:161-231 (score-function estimator; host-composed
        from the same kernels -- a "next" row)."""
        l_signal = self._log_joint_term() + self._entropy_term()
        baseline_cost = None
        if variance_reduction:
            if baseline is not None:
                baseline_cost = 0.5 * torch.square(
                    l_signal.detach() - baseline)
                if self._axis is not None:
                    baseline_cost = ops.reduce_axes(baseline_cost,
                                                    ops.OP_MEAN, self._axis)
                l_signal = l_signal - baseline
            bc = l_signal.detach().mean()
            if not hasattr(self, "_moving_mean"):
                self._moving_mean = torch.zeros((), device=bc.device)
            # assign_moving_average: mm -= (1 - decay) * (mm - bc)
            self._moving_mean = self._moving_mean - (1 - decay) * (
                self._moving_mean - bc)
            l_signal = l_signal - self._moving_mean
        cost = -self._log_joint_term()
        if self._entropy_term() is not None:
            cost = cost + l_signal.detach() * self._entropy_term()
        if self._axis is not None:
            cost = ops.reduce_axes(cost, ops.OP_MEAN, self._axis)
        if baseline_cost is not None:
            return cost, baseline_cost
        return cost


def elbo(meta_bn, observed, latent=None, axis=None, variational=None):
    """exclusive_kl.py:234-267."""
    return EvidenceLowerBoundObjective(
        meta_bn, observed, latent=latent, axis=axis, variational=variational)
