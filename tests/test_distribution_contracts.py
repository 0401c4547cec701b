"""Shape / dtype contract of the whole distribution registry on the CPU (no kernel is
launched): the cases of the reference's shared test helpers, tests/distributions/utils.py:13-520
(`test_batch_shape_2parameter_univariate`, `test_2parameter_sample_shape_same`,
`test_dtype_2parameter`, `test_batch_shape_1parameter`, ...), applied to every class."""
import numpy as np
import pytest
import torch

import zhusuan_b200 as zs

D = zs.distributions

TWO_PARAM = {   # name -> (constructor from two broadcastable float tensors, samples on CPU?)
    "Normal": (lambda a, b: D.Normal(a, std=b.abs() + 1), False),
    # every sampler draws on the device (in-kernel Philox; there is no CPU fallback): the sample
    # shapes of all of these are checked in tests/test_gpu_samplers.py
    "FoldNormal": (lambda a, b: D.FoldNormal(a, std=b.abs() + 1), False),
    "Uniform": (lambda a, b: D.Uniform(a, b.abs() + 1), False),
    # the gamma family samples on the device sampler (csrc/samplers.cu): shapes checked in
    # tests/test_gpu_samplers.py::test_gamma_family_sample_shapes
    "Gamma": (lambda a, b: D.Gamma(a.abs() + 1, b.abs() + 1), False),
    "Beta": (lambda a, b: D.Beta(a.abs() + 1, b.abs() + 1), False),
    "InverseGamma": (lambda a, b: D.InverseGamma(a.abs() + 1, b.abs() + 1), False),
    "Laplace": (lambda a, b: D.Laplace(a, b.abs() + 1), False),
}
BATCH_CASES = [([2, 3], [], [2, 3]), ([2, 3], [3], [2, 3]), ([2, 1, 4], [2, 3, 4], [2, 3, 4]),
               ([2, 3, 5], [3, 1], [2, 3, 5]), ([1, 2, 3], [1, 3], [1, 2, 3])]
SAMPLE_CASES = [([2, 3], [], None, [2, 3]), ([2, 3], [], 1, [1, 2, 3]), ([5], [5], 2, [2, 5]),
                ([2, 1, 4], [1, 2, 4], 3, [3, 2, 2, 4]), ([2, 3], [2, 1], 1, [1, 2, 3]),
                ([1, 3], [], 2, [2, 1, 3]), ([2, 1, 5], [3, 1], 3, [3, 2, 3, 5])]


@pytest.mark.parametrize("name", sorted(TWO_PARAM))
def test_two_parameter_univariate_contract(name):
    make, cpu_sampling = TWO_PARAM[name]
    for s1, s2, target in BATCH_CASES:
        d = make(torch.zeros(s1), torch.ones(s2))
        assert list(d.get_batch_shape()) == target and list(d.batch_shape) == target
        assert list(d.get_value_shape()) == [] and d.dtype == torch.float32
    with pytest.raises(ValueError, match="should be broadcastable to match"):
        make(torch.zeros(2, 3, 5), torch.ones(3, 2))
    if cpu_sampling:
        for s1, s2, n, target in SAMPLE_CASES:
            x = make(torch.zeros(s1), torch.ones(s2)).sample(n)
            assert list(x.shape) == target and x.dtype == torch.float32
    # dtype rules of test_dtype_2parameter: same float dtype or TypeError
    with pytest.raises(TypeError, match="must have the same dtype as"):
        make(torch.zeros(2), torch.ones(2, dtype=torch.float64))
    with pytest.raises(TypeError, match="must have a dtype in"):
        make(torch.zeros(2, dtype=torch.int32), torch.ones(2, dtype=torch.int32))
    d = make(torch.zeros(2), torch.ones(2))
    with pytest.raises(ValueError, match=r"broadcast to match batch_shape \+ value_shape"):
        d.log_prob(torch.zeros(3))
    assert make(torch.zeros(2, 3), torch.ones(3)).group_ndims == 0


ONE_PARAM_DISCRETE = {
    "Bernoulli": lambda p, **kw: D.Bernoulli(p, **kw),
    "Poisson": lambda p, **kw: D.Poisson(p.abs() + 1, **kw),
    "Binomial": lambda p, **kw: D.Binomial(p, 10, **kw),
}


@pytest.mark.parametrize("name", sorted(ONE_PARAM_DISCRETE))
def test_one_parameter_discrete_contract(name):
    """test_batch_shape_1parameter + test_dtype_1parameter_discrete."""
    make = ONE_PARAM_DISCRETE[name]
    for shape in ([], [1], [2, 3, 4]):
        d = make(torch.zeros(shape))
        assert list(d.get_batch_shape()) == shape and list(d.get_value_shape()) == []
        assert d.dtype == torch.int32 and d.param_dtype == torch.float32
        assert d.is_continuous is False and d.is_reparameterized is False
    for dt in (torch.int16, torch.int32, torch.int64, torch.float16, torch.float32, torch.float64):
        assert make(torch.zeros(3), dtype=dt).dtype == dt
    with pytest.raises(TypeError):
        make(torch.zeros(3), dtype=torch.uint8)
    with pytest.raises(TypeError, match="must have a dtype in"):
        make(torch.zeros(3, dtype=torch.int32))
    # (all three draw on device kernels: sample shapes / dtype are checked in
    #  tests/test_gpu_samplers.py::test_count_sample_shapes and test_gpu_distributions.py)


VECTOR_VALUED = {    # [..., n] parameter -> value shape [n], batch shape [...]
    "Categorical": (lambda l: D.Categorical(l), []),
    "OnehotCategorical": (lambda l: D.OnehotCategorical(l), None),
    "Multinomial": (lambda l: D.Multinomial(l, 7), None),
    "UnnormalizedMultinomial": (lambda l: D.UnnormalizedMultinomial(l), None),
    "Dirichlet": (lambda l: D.Dirichlet(l.abs() + 1), None),
    "Concrete": (lambda l: D.Concrete(torch.tensor(1.), l), None),
    "ExpConcrete": (lambda l: D.ExpConcrete(torch.tensor(1.), l), None),
}


@pytest.mark.parametrize("name", sorted(VECTOR_VALUED))
def test_vector_parameter_contract(name):
    """test_batch_shape_1parameter(is_univariate=False) / *_one_rank_less helpers."""
    make, vshape = VECTOR_VALUED[name]
    for shape in ([2], [3, 5], [2, 1, 4]):
        d = make(torch.zeros(shape))
        assert list(d.get_batch_shape()) == shape[:-1]
        assert list(d.get_value_shape()) == ([shape[-1]] if vshape is None else vshape)
    with pytest.raises(ValueError, match="rank"):
        make(torch.zeros([]))
    with pytest.raises(TypeError, match="must have a dtype in"):
        make(torch.zeros(3, dtype=torch.int32))
    # (OnehotCategorical / Multinomial / Dirichlet draw on the device samplers: their sample
    #  shapes are checked in tests/test_gpu_samplers.py)
    if name == "Dirichlet":
        with pytest.raises(ValueError, match="at least 2"):
            make(torch.zeros(3, 1))
    if name == "UnnormalizedMultinomial":
        with pytest.raises(NotImplementedError, match="does not support sampling"):
            make(torch.zeros(3)).sample(1)


def test_matrix_and_multivariate_normal_contract():
    m = D.MultivariateNormalCholesky(torch.zeros(4, 3), torch.eye(3).expand(4, 3, 3))
    assert list(m.get_batch_shape()) == [4] and list(m.get_value_shape()) == [3]
    with pytest.raises(ValueError):
        D.MultivariateNormalCholesky(torch.zeros(4, 3), torch.eye(2).expand(4, 2, 2))
    mv = D.MatrixVariateNormalCholesky(torch.zeros(5, 2, 3), torch.eye(2).expand(5, 2, 2),
                                       torch.eye(3).expand(5, 3, 3))
    assert list(mv.get_batch_shape()) == [5] and list(mv.get_value_shape()) == [2, 3]
    # (sample shapes: tests/test_gpu_samplers.py::test_vector_valued_sample_shapes)
    with pytest.raises(ValueError, match="v_tril should have compatible shape"):
        D.MatrixVariateNormalCholesky(torch.zeros(5, 2, 3), torch.eye(2).expand(5, 2, 2),
                                      torch.eye(2).expand(5, 2, 2))
    with pytest.raises(TypeError, match="must have the same dtype as"):
        D.MatrixVariateNormalCholesky(torch.zeros(2, 3), torch.eye(2).double(), torch.eye(3))
    b = D.BinConcrete(torch.tensor(0.5), torch.zeros(2, 3))
    assert list(b.get_batch_shape()) == [2, 3]      # (sample shape: GPU sampler tests)


def test_group_ndims_validation_everywhere():
    for make in (lambda **k: D.Normal(0., std=1., **k), lambda **k: D.Gamma(torch.ones(2), torch.ones(2), **k),
                 lambda **k: D.Poisson(torch.ones(2), **k), lambda **k: D.Dirichlet(torch.ones(3), **k),
                 lambda **k: D.OnehotCategorical(torch.zeros(3), **k)):
        assert make(group_ndims=torch.tensor(1)).group_ndims == 1
        with pytest.raises(ValueError, match="must be non-negative"):
            make(group_ndims=-1)
        with pytest.raises(ValueError, match="should be a scalar"):
            make(group_ndims=torch.tensor([1, 2]))
        with pytest.raises(ValueError, match="group_event_ndims"):
            make(group_event_ndims=1)
