from .utils import *
from .meta_bn import *
from .bn import *
