"""MetaBayesianNet / meta_bayesian_net (zhusuan/framework/meta_bn.py:29-148)."""
import copy
from functools import wraps

from .utils import Context

__all__ = ["MetaBayesianNet", "meta_bayesian_net"]


class Local(Context):
    def __getattr__(self, item):
        return self.__dict__.get(item, None)

    def __setattr__(self, key, value):
        self.__dict__[key] = value


class MetaBayesianNet(object):
    """A lazily-built Bayesian net: ``observe(**obs)`` re-runs the builder
    under a Local context carrying the observations (meta_bn.py:87-106)."""

    def __init__(self, f, args=None, kwargs=None, scope=None,
                 reuse_variables=False):
        if reuse_variables and scope is None:
            raise ValueError("Cannot reuse tensorflow Variables when `scope` "
                             "is not provided.")
        self._f = f
        self._args = copy.copy(args) or ()
        self._kwargs = copy.copy(kwargs) or {}
        self._scope = scope
        self._reuse_variables = reuse_variables
        self._log_joint = None

    @property
    def log_joint(self):
        return self._log_joint

    @log_joint.setter
    def log_joint(self, value):
        self._log_joint = value

    def _run_with_observations(self, func, observations):
        with Local() as local_cxt:
            local_cxt.observations = observations
            local_cxt.meta_bn = self
            return func(*self._args, **self._kwargs)

    def observe(self, **kwargs):
        return self._run_with_observations(self._f, kwargs)


def meta_bayesian_net(scope=None, reuse_variables=False):
    """Decorator turning a BayesianNet-building function into a
    MetaBayesianNet factory (meta_bn.py:109-148)."""
    def wrapper(f):
        @wraps(f)
        def _wrapped(*args, **kwargs):
            return MetaBayesianNet(f, args=args, kwargs=kwargs, scope=scope,
                                   reuse_variables=reuse_variables)
        return _wrapped
    return wrapper
