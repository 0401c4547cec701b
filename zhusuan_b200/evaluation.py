"""zhusuan/evaluation.py:22-54: importance-sampling marginal likelihood."""
from .variational.monte_carlo import ImportanceWeightedObjective

__all__ = ["is_loglikelihood"]


def is_loglikelihood(meta_bn, observed, latent=None, axis=None, proposal=None):
    """log p(x) >= log_mean_exp_axis(log p(x,z) - log q(z)); identical to
    ``ImportanceWeightedObjective.tensor`` (evaluation.py:50-54)."""
    return ImportanceWeightedObjective(
        meta_bn, observed, latent=latent, axis=axis,
        variational=proposal).tensor
