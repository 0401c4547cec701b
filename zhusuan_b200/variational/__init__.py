"""Variational objectives of the accelerated path: ELBO (exclusive KL), the
importance-weighted bound, the inclusive-KL objective and their gradient
estimators (sgvb / reinforce / vimco / importance)."""
from .base import VariationalObjective
from .exclusive_kl import EvidenceLowerBoundObjective, elbo
from .monte_carlo import (ImportanceWeightedObjective, iw_objective,
                          importance_weighted_objective)
from .inclusive_kl import InclusiveKLObjective, klpq

__all__ = ["VariationalObjective", "EvidenceLowerBoundObjective", "elbo",
           "ImportanceWeightedObjective", "importance_weighted_objective",
           "iw_objective", "InclusiveKLObjective", "klpq"]
