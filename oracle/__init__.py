"""CPU oracle for the zhusuan hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy float32 (and float64 twins where useful), the
arithmetic that thu-ml/zhusuan builds as TensorFlow-1.x graphs for

  * zhusuan/hmc.py              (HMC iteration, step-size tuner, EWMV mass)
  * zhusuan/sgmcmc.py           (SGLD / PSGLD / SGHMC / SGNHT updates)
  * zhusuan/utils.py:177-196    (log_mean_exp)
  * zhusuan/variational/*       (ELBO / IWAE objectives and SGVB gradients)
  * zhusuan/distributions/*     (log_prob of the registry: the six hot-path
                                 distributions and the fourteen others)
  * zhusuan/diagnostics.py      (effective sample size)

Nothing under ``zhusuan_b200/`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or as the timed CPU
baseline -- never as a fallback for the CUDA path.

Pinning status:
  * distribution ``log_prob``: PINNED to the reference's own known-answer
    tests (SciPy formulas and literals from tests/distributions/*.py,
    docs/tutorials/concepts.rst:101-105) -- see tests/test_oracle_pins.py.
  * ``log_mean_exp``: PINNED to tests/test_utils.py:257-284 golden arrays.
  * ELBO / IWAE value and SGVB gradient: PINNED to the analytic-KL checks of
    tests/variational/test_exclusive_kl.py:26-78, test_monte_carlo.py:25-102.
  * VIMCO / REINFORCE / self-normalised importance estimators: PINNED to the
    gradient checks of test_monte_carlo.py:104-142, test_exclusive_kl.py:80-112,
    test_inclusive_kl.py:44-72 (the reference's sample streams and thresholds).
  * the other univariate / multivariate densities: PINNED to every `_test_value`
    literal of tests/distributions/test_univariate.py and test_multivariate.py
    (SciPy targets as the reference computes them), tests/cases.py.
  * effective sample size: PINNED to the properties of tests/test_diagnostics.py.
  * HMC trajectory / accept decision / step-size search / dual averaging / mass
    adaptation and the eight SG-MCMC update rules: PINNED to outputs of the
    reference's own zhusuan/hmc.py and zhusuan/sgmcmc.py, imported unmodified
    and executed on the NumPy stand-in for the TF-1.x graph API in
    oracle/tf_shim/ (TensorFlow itself is not installable here).  Vectors:
    tests/golden/ref_*.npz, written by oracle/tf_shim/make_ref_golden.py;
    checked by tests/test_ref_pins.py (bit-exact on element-wise models,
    float32 matmul-order rounding on dense ones) and regenerated + compared
    whenever /root/reference is present.  The reference's statistical harness
    (tests/test_mcmc.py) is re-run against the oracle as well.
  * ELBO / IWAE .sgvb(), REINFORCE with its moving-mean baseline, VIMCO and the
    inclusive-KL importance estimator: additionally PINNED to the reference's
    own framework/ + distributions/ + variational/ code run on the stand-in
    on the VAE of examples/variational_autoencoders/iwae.py
    (tests/golden/ref_vae.npz: bounds, costs, tf.gradients of every weight).
  * the BNN and LNTM log-joints (oracle/models.py, hand-derived gradients): PINNED to
    bnn_sgmcmc.py:19-35 / lntm_mcem.py:33-48 built on the reference's own
    BayesianNet and sampled by its SGHMC / HMC on the stand-in
    (tests/golden/ref_bnn_sghmc.npz, ref_lntm_hmc.npz).
  * AIS (oracle/evaluation.py): PINNED the same way -- class AIS of
    zhusuan/evaluation.py:57-172 run on the stand-in, tests/golden/ref_ais.npz.
  * device sampler streams (oracle/samplers.py): "parity unpinned" by the
    reference (TF's generators; it tests shapes and moments only); the
    restatement of the published algorithms is the pin there.
"""
