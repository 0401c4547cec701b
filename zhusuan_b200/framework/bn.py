"""StochasticTensor / BayesianNet -- the model-definition contract of
zhusuan/framework/bn.py:26-303, 319-490, 493-1249 (host side; no arithmetic
lives here, every number comes from ``dist.sample`` / ``dist.log_prob``)."""
import warnings

import torch

from .. import distributions
from ..utils import TensorArithmeticMixin, convert_to_tensor
from .meta_bn import Local, MetaBayesianNet
from .utils import Context

__all__ = ["StochasticTensor", "BayesianNet"]


class StochasticTensor(TensorArithmeticMixin):
    """bn.py:26-303.  ``dist`` is duck-typed: it needs ``.dtype``,
    ``.sample(n_samples)``, ``.log_prob(given)``, ``.prob(given)`` (and
    ``get_batch_shape``/``get_value_shape`` when observed) -- the contract the
    reference pins with a Mock in tests/framework/test_base.py:17-40."""

    def __init__(self, bn, name, dist, observation=None, **kwargs):
        if bn is None:
            warnings.warn(
                "The old-style StochasticTensor wrappers will be removed "
                "in a future version. Please see tutorials/concepts.rst for "
                "the suggested way of model construction.", FutureWarning)
            try:
                bn = BayesianNet.get_context()
            except RuntimeError:
                pass
            else:
                bn.nodes[name] = self
        self._bn = bn
        self._name = name
        self._dist = dist
        self._dtype = dist.dtype
        self._n_samples = kwargs.get("n_samples", None)
        if observation is not None:
            self._observation = self._check_observation(observation)
        elif (self._bn is not None) and (self._name in self._bn._observed):
            self._observation = self._check_observation(
                self._bn._observed[name])
        else:
            self._observation = None
        super(StochasticTensor, self).__init__()

    def _check_observation(self, observation):
        type_msg = "Incompatible types of {}('{}') and its observation: {}"
        try:
            if isinstance(observation, TensorArithmeticMixin):
                observation = observation.tensor
            if isinstance(observation, torch.Tensor):
                if isinstance(self._dtype, torch.dtype) and \
                        observation.dtype != self._dtype:
                    if observation.dtype.is_floating_point != \
                            self._dtype.is_floating_point:
                        raise ValueError(
                            "Tensor conversion requested dtype {} for "
                            "Tensor with dtype {}".format(
                                self._dtype, observation.dtype))
                    observation = observation.to(self._dtype)
            else:
                observation = convert_to_tensor(
                    observation, dtype=self._dtype
                    if isinstance(self._dtype, torch.dtype) else None)
        except ValueError as e:
            raise type(e)(type_msg.format(
                self.__class__.__name__, self._name, e))
        shape_msg = "Incompatible shapes of {}('{}') and its observation: " \
                    "{} vs {}."
        try:
            dist_shape = tuple(self._dist.get_batch_shape()) + tuple(
                self._dist.get_value_shape())
        except (AttributeError, TypeError):
            return observation
        try:
            torch.broadcast_shapes(dist_shape, tuple(observation.shape))
        except RuntimeError:
            raise ValueError(shape_msg.format(
                self.__class__.__name__, self._name, dist_shape,
                tuple(observation.shape)))
        return observation

    bn = property(lambda self: self._bn)
    name = property(lambda self: self._name)
    dtype = property(lambda self: self._dtype)
    dist = property(lambda self: self._dist)

    def is_observed(self):
        return self._observation is not None

    @property
    def tensor(self):
        """bn.py:163-175: the observation, else a (cached) sample."""
        if self._observation is not None:
            return self._observation
        elif not hasattr(self, "_samples"):
            self._samples = self._dist.sample(n_samples=self._n_samples)
        return self._samples

    @property
    def shape(self):
        return self.tensor.shape

    def get_shape(self):
        return self.shape

    @property
    def cond_log_p(self):
        """bn.py:194-204: log p(value | parents), cached."""
        if not hasattr(self, "_cond_log_p"):
            self._cond_log_p = self._dist.log_prob(self.tensor)
        return self._cond_log_p

    # deprecated surface kept for drop-in parity (bn.py:216-303)
    @property
    def net(self):
        warnings.warn("StochasticTensor: The `.net` property will be removed "
                      "in the coming version (0.4.1), use `.bn` instead.",
                      FutureWarning)
        return self._bn

    @property
    def distribution(self):
        warnings.warn("StochasticTensor: The `.distribution` property will be "
                      "removed in the coming version (0.4.1), use `.dist` "
                      "instead.", FutureWarning)
        return self._dist

    def sample(self, n_samples):
        warnings.warn("StochasticTensor: The `sample()` method will be "
                      "removed in the coming version (0.4.1), use "
                      "`.dist.sample()` instead.", FutureWarning)
        return self._dist.sample(n_samples)

    def log_prob(self, given):
        warnings.warn("StochasticTensor: The `log_prob()` method will be "
                      "removed in the coming version (0.4.1), use "
                      "`.dist.log_prob()` instead.", FutureWarning)
        return self._dist.log_prob(given)

    def prob(self, given):
        warnings.warn("StochasticTensor: The `prob()` method will be removed "
                      "in the coming version (0.4.1), use `.dist.prob()` "
                      "instead.", FutureWarning)
        return self._dist.prob(given)


class _BayesianNet(object):
    """bn.py:319-490."""

    def __init__(self):
        self._nodes = {}
        try:
            self._local_cxt = Local.get_context()
        except RuntimeError:
            self._local_cxt = None
        self._meta_bn = self._local_cxt.meta_bn if self._local_cxt else None
        super(_BayesianNet, self).__init__()

    nodes = property(lambda self: self._nodes)

    def _get_observation(self, name):
        if self._local_cxt:
            return self._local_cxt.observations.get(name, None)
        return None

    def stochastic(self, name, dist, **kwargs):
        if name in self._nodes:
            raise ValueError(
                "There exists a node with name '{}' in the {}. Names should "
                "be unique.".format(name, BayesianNet.__name__))
        if hasattr(self, "_log_joint_cache"):
            del self._log_joint_cache
        node = StochasticTensor(
            self, name, dist, observation=self._get_observation(name),
            **kwargs)
        self._nodes[name] = node
        return node

    def deterministic(self, name, input_tensor):
        input_tensor = convert_to_tensor(input_tensor)
        self._nodes[name] = input_tensor
        return input_tensor

    def _check_name_exist(self, name, only_stochastic=False):
        if not isinstance(name, str):
            raise TypeError(
                "Expected string in `name_or_names`, got {} of type {}."
                .format(repr(name), type(name)))
        if name not in self._nodes:
            raise ValueError("There isn't a node named '{}' in the {}."
                             .format(name, BayesianNet.__name__))
        elif only_stochastic and not isinstance(
                self._nodes[name], StochasticTensor):
            raise ValueError("Node '{}' is deterministic (input or output)."
                             .format(name))
        return name

    def _check_names_exist(self, name_or_names, only_stochastic=False):
        if isinstance(name_or_names, str):
            names = (name_or_names,)
        else:
            name_or_names = tuple(name_or_names)
            names = name_or_names
        for name in names:
            self._check_name_exist(name, only_stochastic=only_stochastic)
        return name_or_names

    def get(self, name_or_names):
        name_or_names = self._check_names_exist(name_or_names)
        if isinstance(name_or_names, tuple):
            return [self._nodes[name] for name in name_or_names]
        return self._nodes[name_or_names]

    def cond_log_prob(self, name_or_names):
        name_or_names = self._check_names_exist(name_or_names,
                                                only_stochastic=True)
        if isinstance(name_or_names, tuple):
            return [self._nodes[name].cond_log_p for name in name_or_names]
        return self._nodes[name_or_names].cond_log_p

    def _log_joint(self):
        """bn.py:454-465: sum of cond_log_p, or the meta_bn's override."""
        if (self._meta_bn is None) or (self._meta_bn.log_joint is None):
            return sum(node.cond_log_p for node in self._nodes.values()
                       if isinstance(node, StochasticTensor))
        elif callable(self._meta_bn.log_joint):
            return self._meta_bn.log_joint(self)
        raise TypeError(
            "{}.log_joint is set to a non-callable instance: {}".format(
                self._meta_bn.__class__.__name__,
                repr(self._meta_bn.log_joint)))

    def log_joint(self):
        if not hasattr(self, "_log_joint_cache"):
            self._log_joint_cache = self._log_joint()
        return self._log_joint_cache

    def __getitem__(self, name):
        return self._nodes[self._check_name_exist(name)]

    def __setitem__(self, name, node):
        raise TypeError(
            "{} instance does not support replacement of the existing node. "
            "To achieve this, pass observations of certain nodes when "
            "calling {}.{}".format(
                BayesianNet.__name__, MetaBayesianNet.__name__,
                MetaBayesianNet.observe.__name__))


class BayesianNet(_BayesianNet, Context):
    """bn.py:493-1249 -- one factory method per registry distribution."""

    def __init__(self, observed=None):
        self._observed = observed if observed else {}
        super(BayesianNet, self).__init__()

    # ---- the 0.3-era query API, kept (deprecated) by the reference: bn.py:1200-1249
    def outputs(self, name_or_names):
        warnings.warn(
            "BayesianNet: `outputs()` has been deprecated in 0.4 and will "
            "be removed in 0.4.1, use `get()` instead.", FutureWarning)
        nodes = self.get(name_or_names)
        if isinstance(nodes, list):
            return [getattr(n, "tensor", n) for n in nodes]
        return getattr(nodes, "tensor", nodes)

    def local_log_prob(self, name_or_names):
        warnings.warn(
            "BayesianNet: `local_log_prob()` has been deprecated in 0.4 "
            "and will be removed in 0.4.1, use `cond_log_prob()` instead.",
            FutureWarning)
        return self.cond_log_prob(name_or_names)

    def query(self, name_or_names, outputs=False, local_log_prob=False):
        warnings.warn(
            "BayesianNet: `query()` has been deprecated in 0.4 "
            "and will be removed in 0.4.1, use `get()` and "
            "`cond_log_prob()` instead.", FutureWarning)
        ret = []
        if outputs:
            ret.append(self.outputs(name_or_names))
        if local_log_prob:
            ret.append(self.local_log_prob(name_or_names))
        if not ret:
            raise ValueError("No query options are selected.")
        if not isinstance(name_or_names, str):
            return list(zip(*ret))
        return tuple(ret)

    def normal(self, name, mean=0., _sentinel=None, std=None, logstd=None,
               group_ndims=0, n_samples=None, is_reparameterized=True,
               check_numerics=False, **kwargs):
        """bn.py:556-590."""
        dist = distributions.Normal(
            mean, _sentinel=_sentinel, std=std, logstd=logstd,
            group_ndims=group_ndims, is_reparameterized=is_reparameterized,
            check_numerics=check_numerics, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    def bernoulli(self, name, logits, n_samples=None, group_ndims=0,
                  dtype=torch.int32, **kwargs):
        """bn.py:628-654."""
        dist = distributions.Bernoulli(logits, group_ndims=group_ndims,
                                       dtype=dtype, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    def categorical(self, name, logits, n_samples=None, group_ndims=0,
                    dtype=torch.int32, **kwargs):
        """bn.py:656-682."""
        dist = distributions.Categorical(logits, group_ndims=group_ndims,
                                         dtype=dtype, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    discrete = categorical

    def dirichlet(self, name, alpha, n_samples=None, group_ndims=0,
                  check_numerics=False, **kwargs):
        """bn.py:938-965."""
        dist = distributions.Dirichlet(alpha, group_ndims=group_ndims,
                                       check_numerics=check_numerics,
                                       **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    # ---- the other elementwise univariate families (bn.py:592-626, 686-838,
    # 1027-1121); one generic factory body, the reference's signatures
    def _univariate(self, cls, name, args, n_samples, kwargs, **dist_kw):
        dist = cls(*args, **dict(dist_kw, **kwargs))
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    def fold_normal(self, name, mean=0., _sentinel=None, std=None, logstd=None,
                    n_samples=None, group_ndims=0, is_reparameterized=True,
                    check_numerics=False, **kwargs):
        return self._univariate(
            distributions.FoldNormal, name, (mean,), n_samples, kwargs,
            _sentinel=_sentinel, std=std, logstd=logstd,
            group_ndims=group_ndims, is_reparameterized=is_reparameterized,
            check_numerics=check_numerics)

    def uniform(self, name, minval=0., maxval=1., n_samples=None,
                group_ndims=0, is_reparameterized=True, check_numerics=False,
                **kwargs):
        return self._univariate(
            distributions.Uniform, name, (minval, maxval), n_samples, kwargs,
            group_ndims=group_ndims, is_reparameterized=is_reparameterized,
            check_numerics=check_numerics)

    def gamma(self, name, alpha, beta, n_samples=None, group_ndims=0,
              check_numerics=False, **kwargs):
        return self._univariate(
            distributions.Gamma, name, (alpha, beta), n_samples, kwargs,
            group_ndims=group_ndims, check_numerics=check_numerics)

    def beta(self, name, alpha, beta, n_samples=None, group_ndims=0,
             check_numerics=False, **kwargs):
        return self._univariate(
            distributions.Beta, name, (alpha, beta), n_samples, kwargs,
            group_ndims=group_ndims, check_numerics=check_numerics)

    def inverse_gamma(self, name, alpha, beta, n_samples=None, group_ndims=0,
                      check_numerics=False, **kwargs):
        return self._univariate(
            distributions.InverseGamma, name, (alpha, beta), n_samples, kwargs,
            group_ndims=group_ndims, check_numerics=check_numerics)

    def poisson(self, name, rate, n_samples=None, group_ndims=0,
                dtype=torch.int32, check_numerics=False, **kwargs):
        return self._univariate(
            distributions.Poisson, name, (rate,), n_samples, kwargs,
            group_ndims=group_ndims, dtype=dtype,
            check_numerics=check_numerics)

    def binomial(self, name, logits, n_experiments, n_samples=None,
                 group_ndims=0, dtype=torch.int32, check_numerics=False,
                 **kwargs):
        return self._univariate(
            distributions.Binomial, name, (logits, n_experiments), n_samples,
            kwargs, group_ndims=group_ndims, dtype=dtype,
            check_numerics=check_numerics)

    def laplace(self, name, loc, scale, n_samples=None, group_ndims=0,
                is_reparameterized=True, check_numerics=False, **kwargs):
        return self._univariate(
            distributions.Laplace, name, (loc, scale), n_samples, kwargs,
            group_ndims=group_ndims, is_reparameterized=is_reparameterized,
            check_numerics=check_numerics)

    def bin_concrete(self, name, temperature, logits, n_samples=None,
                     group_ndims=0, is_reparameterized=True,
                     check_numerics=False, **kwargs):
        return self._univariate(
            distributions.BinConcrete, name, (temperature, logits), n_samples,
            kwargs, group_ndims=group_ndims,
            is_reparameterized=is_reparameterized,
            check_numerics=check_numerics)

    bin_gumbel_softmax = bin_concrete

    def unnormalized_multinomial(self, name, logits, normalize_logits=True,
                                 group_ndims=0, dtype=torch.int32, **kwargs):
        """bn.py:938-965 (no ``n_samples``: the distribution cannot sample)."""
        dist = distributions.UnnormalizedMultinomial(
            logits, normalize_logits=normalize_logits,
            group_ndims=group_ndims, dtype=dtype, **kwargs)
        return self.stochastic(name, dist, **kwargs)

    bag_of_categoricals = unnormalized_multinomial

    def multinomial(self, name, logits, n_experiments, normalize_logits=True,
                    n_samples=None, group_ndims=0, dtype=torch.int32,
                    **kwargs):
        """bn.py:872-904."""
        dist = distributions.Multinomial(
            logits, n_experiments, normalize_logits=normalize_logits,
            group_ndims=group_ndims, dtype=dtype, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    def onehot_categorical(self, name, logits, n_samples=None, group_ndims=0,
                           dtype=torch.int32, **kwargs):
        """bn.py:906-934."""
        dist = distributions.OnehotCategorical(
            logits, group_ndims=group_ndims, dtype=dtype, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    onehot_discrete = onehot_categorical

    def exp_concrete(self, name, temperature, logits, n_samples=None,
                     group_ndims=0, is_reparameterized=True,
                     check_numerics=False, **kwargs):
        """bn.py:1123-1155."""
        dist = distributions.ExpConcrete(
            temperature, logits, group_ndims=group_ndims,
            is_reparameterized=is_reparameterized,
            check_numerics=check_numerics, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    exp_gumbel_softmax = exp_concrete

    def concrete(self, name, temperature, logits, n_samples=None,
                 group_ndims=0, is_reparameterized=True, check_numerics=False,
                 **kwargs):
        """bn.py:1157-1189."""
        dist = distributions.Concrete(
            temperature, logits, group_ndims=group_ndims,
            is_reparameterized=is_reparameterized,
            check_numerics=check_numerics, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    gumbel_softmax = concrete

    def matrix_variate_normal_cholesky(self, name, mean, u_tril, v_tril,
                                       n_samples=None, group_ndims=0,
                                       is_reparameterized=True,
                                       check_numerics=False, **kwargs):
        """bn.py:967-997."""
        dist = distributions.MatrixVariateNormalCholesky(
            mean, u_tril, v_tril, group_ndims=group_ndims,
            is_reparameterized=is_reparameterized,
            check_numerics=check_numerics, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)

    def multivariate_normal_cholesky(self, name, mean, cov_tril,
                                     n_samples=None, group_ndims=0,
                                     is_reparameterized=True,
                                     check_numerics=False, **kwargs):
        """bn.py:999-1025."""
        dist = distributions.MultivariateNormalCholesky(
            mean, cov_tril, group_ndims=group_ndims,
            is_reparameterized=is_reparameterized,
            check_numerics=check_numerics, **kwargs)
        return self.stochastic(name, dist, n_samples=n_samples, **kwargs)
