"""Distribution registry: the 20 distributions of zhusuan/distributions (plus
their aliases); ``log_prob`` of each runs on the libzsb200 kernels."""
from .base import Distribution
from .univariate import (Normal, Bernoulli, Categorical, Discrete)
from .univariate_more import (FoldNormal, Uniform, Gamma, Beta, Poisson, Binomial, InverseGamma, Laplace, BinConcrete, BinGumbelSoftmax)
from .multivariate import (MultivariateNormalCholesky, UnnormalizedMultinomial, BagofCategoricals, Dirichlet, Multinomial, OnehotCategorical, OnehotDiscrete)
from .multivariate_more import (ExpConcrete, ExpGumbelSoftmax, Concrete, GumbelSoftmax, MatrixVariateNormalCholesky)

__all__ = ['Distribution', 'Normal', 'Bernoulli', 'Categorical', 'Discrete', 'FoldNormal', 'Uniform', 'Gamma', 'Beta', 'Poisson', 'Binomial', 'InverseGamma', 'Laplace', 'BinConcrete', 'BinGumbelSoftmax', 'MultivariateNormalCholesky', 'UnnormalizedMultinomial', 'BagofCategoricals', 'Dirichlet', 'Multinomial', 'OnehotCategorical', 'OnehotDiscrete', 'ExpConcrete', 'ExpGumbelSoftmax', 'Concrete', 'GumbelSoftmax', 'MatrixVariateNormalCholesky']
