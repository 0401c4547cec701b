from .base import *
from .univariate import *
from .multivariate import *
