"""VariationalObjective (zhusuan/variational/base.py:24-196): wiring between a
model (MetaBayesianNet or log-joint callable), the variational family's
samples and their log q(z); Tensor-like (``.tensor``, arithmetic)."""
import copy
import warnings

from ..framework.bn import StochasticTensor, BayesianNet
from ..utils import TensorArithmeticMixin, merge_dicts

__all__ = ['VariationalObjective']


class VariationalObjective(TensorArithmeticMixin):
    def __init__(self, meta_bn, observed, latent=None, variational=None):
        if callable(meta_bn):
            self._meta_bn = None
            self._log_joint = meta_bn
        else:
            self._meta_bn = meta_bn
        if (variational is None) == (latent is None):      # base.py:56-62
            raise ValueError(
                "Either a {} `variational` representing "
                "the variational family or a dictionary `latent` "
                "representing the variational inputs should be passed. "
                "It is not allowed that both are specified or both are not."
                .format(BayesianNet))
        elif latent is None:
            if isinstance(variational, BayesianNet):
                self._variational = variational
            else:                                          # base.py:66-69
                raise TypeError(
                    "`variational` should be a {} instance, got {}."
                    .format(BayesianNet.__name__, repr(variational)))
            v_inputs = [(name, node)
                        for name, node in self._variational.nodes.items()
                        if isinstance(node, StochasticTensor) and
                        not node.is_observed()]            # base.py:70-72
            v_log_probs = [(name, node.cond_log_p) for name, node in v_inputs]
        else:
            warnings.warn(
                "The `latent` argument has been deprecated and will be "
                "removed in the coming version (0.4.1), use the `variational` "
                "argument instead.", FutureWarning)
            self._variational = None
            v_inputs = [(k, v[0]) for k, v in latent.items()]
            v_log_probs = [(k, v[1]) for k, v in latent.items()]
        self._v_inputs = dict(v_inputs)
        self._v_log_probs = dict(v_log_probs)
        self._observed = copy.copy(observed)

    def _validate_variational_inputs(self, bn):
        for node in bn.nodes.values():
            if isinstance(node, StochasticTensor) and \
                    (not node.is_observed()):
                raise ValueError(
                    "Stochastic node '{}' in the model is neither "
                    "observed nor provided with a variational posterior."
                    .format(node.name))

    meta_bn = property(lambda self: self._meta_bn)
    variational = property(lambda self: self._variational)

    @property
    def bn(self):
        """base.py:117-138: the model observed at the variational samples."""
        if self._meta_bn:
            if not hasattr(self, "_bn"):
                self._bn = self._meta_bn.observe(
                    **merge_dicts(self._v_inputs, self._observed))
                self._validate_variational_inputs(self._bn)
            return self._bn
        return None

    def _objective(self):
        raise NotImplementedError()

    @property
    def tensor(self):
        if not hasattr(self, "_tensor"):
            self._tensor = self._objective()
        return self._tensor

    def _log_joint_term(self):                             # base.py:169-175
        if self._meta_bn:
            return self.bn.log_joint()
        elif not hasattr(self, '_log_joint_cache'):
            self._log_joint_cache = self._log_joint(
                merge_dicts(self._v_inputs, self._observed))
        return self._log_joint_cache

    def _entropy_term(self):                               # base.py:177-183
        if not hasattr(self, '_entropy_cache'):
            if len(self._v_log_probs) > 0:
                self._entropy_cache = -sum(self._v_log_probs.values())
            else:
                self._entropy_cache = None
        return self._entropy_cache
