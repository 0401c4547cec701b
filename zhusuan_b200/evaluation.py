"""zhusuan/evaluation.py: importance-sampling and annealed-importance-sampling
marginal-likelihood estimators on the B200 kernels."""
import numpy as np
import torch

from .utils import merge_dicts
from .variational.monte_carlo import ImportanceWeightedObjective

__all__ = ["is_loglikelihood", "AIS"]


def is_loglikelihood(meta_bn, observed, latent=None, axis=None, proposal=None):
    """log p(x) >= log_mean_exp_axis(log p(x,z) - log q(z)); identical to
    ``ImportanceWeightedObjective.tensor`` (evaluation.py:22-54)."""
    return ImportanceWeightedObjective(
        meta_bn, observed, latent=latent, axis=axis,
        variational=proposal).tensor


class AIS(object):
    """Annealed importance sampling (evaluation.py:57-172): a host loop over
    ``n_temperatures`` tempered HMC iterations.  Same schedule, adaptation
    phase and weight recursion as the reference; the log-weights accumulate
    on the device (no per-temperature read-back) and only the final bound is
    copied to the host.  ``temperature`` is the reference's placeholder: a
    plain attribute read by the tempered log-joint at every evaluation.
    """

    def __init__(self, meta_bn, proposal_meta_bn, hmc, observed, latent,
                 n_temperatures=1000, n_adapt=30, verbose=False):
        self._n_temperatures = n_temperatures
        self._n_adapt = n_adapt
        self._verbose = verbose
        if callable(meta_bn):
            log_joint = meta_bn
        else:
            log_joint = lambda obs: meta_bn.observe(**obs).log_joint()
        self._latent_k = list(latent.keys())
        self._latent_v = [latent[k] for k in self._latent_k]
        self._proposal = proposal_meta_bn
        log_prior = lambda obs: proposal_meta_bn.observe(**obs).log_joint()
        self.temperature = 0.0

        def log_fn(obs):                                  # evaluation.py:91-94
            t = self.temperature
            return log_prior(obs) * (1 - t) + log_joint(obs) * t
        self.log_fn = log_fn
        self._observed = dict(observed)
        self.sample_op, self.hmc_info = hmc.sample(log_fn, observed, latent)

    def _init_latent(self):
        """evaluation.py:87, 100-101: z <- a fresh sample of the proposal."""
        samples = self._proposal.observe().get(self._latent_k)   # as the reference
        for z, s in zip(self._latent_v, samples):
            z.copy_(torch.as_tensor(s.tensor if hasattr(s, "tensor") else s))

    def _map_t(self, t):
        return 1. / (1. + np.exp(-4 * (2 * t / self._n_temperatures - 1)))

    def _get_schedule_t(self, t):
        return (self._map_t(t) - self._map_t(0)) / (
            self._map_t(self._n_temperatures) - self._map_t(0))

    def run(self, sess=None, feed_dict=None):
        """evaluation.py:119-165.  ``sess`` / ``feed_dict`` are accepted for
        call-compatibility; observed tensors given in ``feed_dict`` by name
        replace the construction-time observations."""
        if feed_dict:
            self._observed.update({k: v for k, v in feed_dict.items()
                                   if isinstance(k, str)})
        adp_num_t = 2 if self._n_temperatures > 1 else 1
        adp_t = self._get_schedule_t(adp_num_t)
        self._init_latent()
        for i in range(self._n_adapt):
            self.temperature = adp_t
            self.sample_op()
            if self._verbose:
                print('Adapt iter {}, acc = {:.3f}'.format(
                    i, float(self.hmc_info.acceptance_rate.mean())))
        self._init_latent()
        self.temperature = 0.0
        with torch.no_grad():
            prior_density = self.log_fn(merge_dicts(
                self._observed, dict(zip(self._latent_k, self._latent_v))))
        log_weights = -prior_density.clone()
        for num_t in range(self._n_temperatures):
            self.temperature = self._get_schedule_t(num_t + 1)
            self.sample_op()
            old_log_p = self.hmc_info.orig_log_prob
            new_log_p = self.hmc_info.log_prob
            if num_t + 1 < self._n_temperatures:
                log_weights += old_log_p - new_log_p
            else:
                log_weights += old_log_p
            if self._verbose:
                print('Finished step {}, Temperature = {:.4f}, acc = {:.3f}'
                      .format(num_t + 1, self.temperature,
                              float(self.hmc_info.acceptance_rate.mean())))
        self.log_weights = log_weights
        return float(self._get_lower_bound(log_weights).mean())

    @staticmethod
    def _get_lower_bound(log_weights):
        """evaluation.py:167-172: log_mean_exp over the chain axis (axis 0)."""
        from . import ops
        return ops.reduce_axes(log_weights.contiguous(), ops.OP_LME, 0)
