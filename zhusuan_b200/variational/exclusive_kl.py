"""ELBO objective (zhusuan/variational/exclusive_kl.py:20-267): ``.tensor``
and ``.sgvb()`` on the K6 ``mean`` reduction kernel (fwd + 1/K backward)."""
import weakref

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

    def reinforce(self, variance_reduction=True, baseline=None, decay=0.8,
                  moving_mean=None):
        """exclusive_kl.py:161-231 (score-function estimator; host-composed
        from the same kernels -- a "next" row).

        The reference keeps the moving-mean baseline in ``tf.get_variable('moving_mean')``
        (exclusive_kl.py:209-216): ONE variable that persists across ``sess.run`` steps.  Here an
        objective is rebuilt every step, so the variable lives in a module-level registry
        (one per device, ``reset_moving_mean()`` clears it) or in the 0-d tensor passed as
        ``moving_mean`` (updated in place) -- not on the short-lived objective instance.

        Update rule = ``moving_averages.assign_moving_average(moving_mean, bc, decay)`` of
        TensorFlow 1.x, whose DEFAULT is ``zero_debias=True``: ``biased -= (biased - bc)(1 - decay)``,
        ``step += 1``, ``moving_mean = biased / (1 - decay**step)`` (so the first update sets it
        to ``bc`` itself).  The learning signal of a step subtracts the value the variable had
        BEFORE that step's update: in the reference graph the read in ``l_signal - moving_mean``
        carries no dependency on ``update_mean`` (exclusive_kl.py:215-219) and is ready first --
        the order in which the reference's own code runs on the NumPy TF stand-in, pinned by
        tests/golden/ref_vae.npz."""
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
            mm = moving_mean
            if mm is None:
                mm = _MOVING_MEAN.get(bc.device)
                if mm is None:
                    mm = _MOVING_MEAN[bc.device] = torch.zeros((), device=bc.device)
            shadow = _SHADOW.get(mm)
            if shadow is None:           # the 'biased' and 'local_step' shadow variables
                shadow = _SHADOW[mm] = [torch.zeros_like(mm), 0]
            l_signal = l_signal - mm.clone()                   # value before this step's update
            shadow[0].sub_((1 - decay) * (shadow[0] - bc))
            shadow[1] += 1
            mm.copy_(shadow[0] / (1 - decay ** shadow[1]))
            self._moving_mean = mm
        cost = -self._log_joint_term()
        if self._entropy_term() is not None:
            cost = cost + l_signal.detach() * self._entropy_term()
        if self._axis is not None:
            cost = ops.reduce_axes(cost, ops.OP_MEAN, self._axis)
        if baseline_cost is not None:
            return cost, baseline_cost
        return cost


_MOVING_MEAN = {}      # the 'moving_mean' variable of exclusive_kl.py:209-212, one per device
_SHADOW = weakref.WeakKeyDictionary()   # moving_mean tensor -> [biased, local_step] (zero_debias)


def reset_moving_mean():
    """Forget the REINFORCE moving-mean baseline (a fresh ``tf.global_variables_initializer``)."""
    for mm in _MOVING_MEAN.values():
        _SHADOW.pop(mm, None)
    _MOVING_MEAN.clear()


def elbo(meta_bn, observed, latent=None, axis=None, variational=None):
    """exclusive_kl.py:234-267."""
    return EvidenceLowerBoundObjective(
        meta_bn, observed, latent=latent, axis=axis, variational=variational)
