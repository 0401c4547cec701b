"""Deferred model construction: ``MetaBayesianNet`` / ``@meta_bayesian_net``.

Contract mirrored from zhusuan/framework/meta_bn.py:29-148: a model is a *function* that builds a
``BayesianNet``; the decorator captures the call (function + arguments) instead of running it, and
every ``observe(**{name: value})`` replays the call with the observations published to the nodes
being built, so that ``bn.stochastic(name, ...)`` picks its value up by name (bn.py:348-371).
``meta_bn.log_joint = fn`` overrides the default sum of conditional log-probs (bn.py:454-465).

How the observations reach the net under construction: a build *frame* (who is building, with
which observations) is pushed on a per-class stack for the duration of the builder call;
``_BayesianNet.__init__`` looks at the innermost frame.  ``Local`` is that frame type -- the name
the reference uses and the one bn.py imports.
"""
import functools

from .utils import Context

__all__ = ["MetaBayesianNet", "meta_bayesian_net"]


class Local(Context):
    """One build frame: ``observations`` (dict name -> value) and ``meta_bn`` (the owner)."""

    def __init__(self, meta_bn=None, observations=None):
        self.meta_bn = meta_bn
        self.observations = dict(observations or {})


class _DeferredCall(object):
    """``fn(*args, **kwargs)`` frozen for later replays (shallow copies, as the reference takes)."""

    __slots__ = ("fn", "args", "kwargs")

    def __init__(self, fn, args, kwargs):
        self.fn = fn
        self.args = tuple(args or ())
        self.kwargs = dict(kwargs or {})

    def __call__(self):
        return self.fn(*self.args, **self.kwargs)


class MetaBayesianNet(object):
    """A Bayesian net that is (re)built on demand.

    :param f: the builder; must return a ``BayesianNet``.
    :param args / kwargs: its arguments (meta_bn.py:50-56).
    :param scope / reuse_variables: TF variable-scope options of the reference.  Parameters
        here are explicit tensors owned by the caller, so they only keep their validation:
        ``reuse_variables`` without a ``scope`` is an error (meta_bn.py:57-60).
    """

    def __init__(self, f, args=None, kwargs=None, scope=None, reuse_variables=False):
        if reuse_variables and scope is None:
            raise ValueError("Cannot reuse tensorflow Variables when `scope` is not provided.")
        self._build = _DeferredCall(f, args, kwargs)
        self._scope, self._reuse_variables = scope, bool(reuse_variables)
        self.log_joint = None       # None: sum of cond_log_p; else a callable bn -> log joint

    def observe(self, **observations):
        """Build the net with the named stochastic nodes fixed to the given values
        (meta_bn.py:87-106) and return it."""
        with Local(meta_bn=self, observations=observations):
            return self._build()


def meta_bayesian_net(scope=None, reuse_variables=False):
    """``@zs.meta_bayesian_net(scope=..., reuse_variables=...)``: calling the decorated function
    no longer builds the net but returns the ``MetaBayesianNet`` holding that call
    (meta_bn.py:109-148)."""
    def decorate(builder):
        @functools.wraps(builder)
        def make(*args, **kwargs):
            return MetaBayesianNet(builder, args=args, kwargs=kwargs, scope=scope,
                                   reuse_variables=reuse_variables)
        return make
    return decorate
