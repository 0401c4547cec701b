from .base import *
from .univariate import *
from .multivariate import *
from .univariate_more import *
from .multivariate_more import *
