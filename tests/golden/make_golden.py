"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).

The reference holds NO golden vectors for an HMC trajectory, accept decision,
step-size / mass adaptation or any SG-MCMC update (tests/test_mcmc.py is
statistical only) and TensorFlow cannot be installed here, so these vectors
are oracle-generated ("parity unpinned" by the reference, see
oracle/__init__.py).  They freeze the oracle: tests/test_oracle_hmc.py checks
the oracle still reproduces them, and the GPU parity tests replay the same
injected noise through the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hmc as OH            # noqa: E402
from oracle import sgmcmc as OS         # noqa: E402
from oracle import models as OM         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_hmc(model, q0, n_iters, n_adapt, rng, **kw):
    h = OH.HMC(**kw)
    q = [q0.copy()]
    rec = {k: [] for k in ("noise_p", "noise_u", "q", "acc", "accept",
                           "step_size", "eps_used", "mass", "lp", "h0", "h1",
                           "lp0", "p0")}
    for i in range(n_iters):
        npz = rng.standard_normal(q0.shape).astype(np.float32)
        nu = rng.random(q0.shape[0]).astype(np.float32)
        adapt = i < n_adapt
        q, info = h.step(q, model.logp, model.grad, [npz], nu,
                         adapt_step_size=adapt, adapt_mass=adapt)
        rec["noise_p"].append(npz)
        rec["noise_u"].append(nu)
        rec["q"].append(q[0].copy())
        rec["acc"].append(info.acceptance_rate)
        rec["accept"].append(info.if_accept.astype(np.int32))
        rec["step_size"].append(np.float32(info.updated_step_size))
        rec["eps_used"].append(np.float32(info.step_size_used))
        rec["mass"].append(info.mass[0].reshape(-1))
        rec["lp"].append(info.log_prob)
        rec["h0"].append(info.orig_hamiltonian)
        rec["h1"].append(info.hamiltonian)
        rec["lp0"].append(info.orig_log_prob)
        rec["p0"].append(info.init_momentum[0])
    out = {k: np.stack(v) for k, v in rec.items()}
    out["n_search_iters"] = np.int32(h.n_search_iters)
    return out


def make_hmc_diag():
    rng = np.random.Generator(np.random.PCG64(101))
    D, C = 12, 16
    std = (1.0 / (1.0 + np.arange(D))).astype(np.float32)   # gaussian.py:29
    model = OM.DiagGaussian(np.zeros(D, np.float32), std)
    q0 = (0.1 * rng.standard_normal((C, D))).astype(np.float32)
    cfg = dict(step_size=1e-3, n_leapfrogs=5, adapt_step_size=True,
               target_acceptance_rate=0.9, adapt_mass=True,
               mass_collect_iters=4, mass_decay=0.99)
    out = run_hmc(model, q0, n_iters=14, n_adapt=9, rng=rng, **cfg)
    out.update(q0=q0, std=std, n_adapt=np.int32(9),
               **{"cfg_" + k: np.float32(v) for k, v in cfg.items()})
    np.savez_compressed(os.path.join(HERE, "hmc_diag.npz"), **out)


def make_hmc_dense():
    rng = np.random.Generator(np.random.PCG64(202))
    D, C = 32, 24
    P, const = OM.make_dense_gaussian_problem(D, seed=2)
    mu = (0.5 * rng.standard_normal(D)).astype(np.float32)
    model = OM.DenseGaussian(P.astype(np.float32), mu, const)
    q0 = rng.standard_normal((C, D)).astype(np.float32)
    cfg = dict(step_size=0.05, n_leapfrogs=4, adapt_step_size=True,
               target_acceptance_rate=0.8, adapt_mass=True,
               mass_collect_iters=3, mass_decay=0.99)
    out = run_hmc(model, q0, n_iters=12, n_adapt=10, rng=rng, **cfg)
    out.update(q0=q0, P=P, mu=mu, const=np.float64(const), n_adapt=np.int32(10),
               **{"cfg_" + k: np.float32(v) for k, v in cfg.items()})
    np.savez_compressed(os.path.join(HERE, "hmc_dense.npz"), **out)


def make_sgmcmc():
    rng = np.random.Generator(np.random.PCG64(303))
    D, C, T = 8, 6, 5
    std = (0.5 + 0.1 * np.arange(D)).astype(np.float32)
    model = OM.DiagGaussian(np.linspace(-1, 1, D).astype(np.float32), std)
    q0 = rng.standard_normal((C, D)).astype(np.float32)
    nz = lambda: rng.standard_normal((C, D)).astype(np.float32)
    out = {"q0": q0, "std": std, "mean": model.mean}
    samplers = {
        "sgld": (OS.SGLD, dict(learning_rate=0.01)),
        "psgld": (OS.PSGLD, dict(learning_rate=0.01)),
        "sghmc1": (OS.SGHMC, dict(learning_rate=0.01, friction=0.3,
                                  variance_estimate=0.02,
                                  n_iter_resample_v=3, second_order=False)),
        "sghmc2": (OS.SGHMC, dict(learning_rate=0.01, friction=0.3,
                                  variance_estimate=0.02,
                                  n_iter_resample_v=3, second_order=True)),
        "sgnht1v": (OS.SGNHT, dict(learning_rate=0.01, variance_extra=0.1,
                                   tune_rate=2., n_iter_resample_v=4,
                                   second_order=False, use_vector_alpha=True)),
        "sgnht2v": (OS.SGNHT, dict(learning_rate=0.01, variance_extra=0.1,
                                   tune_rate=2., n_iter_resample_v=4,
                                   second_order=True, use_vector_alpha=True)),
        "sgnht1s": (OS.SGNHT, dict(learning_rate=0.01, variance_extra=0.1,
                                   tune_rate=2., n_iter_resample_v=None,
                                   second_order=False,
                                   use_vector_alpha=False)),
        "sgnht2s": (OS.SGNHT, dict(learning_rate=0.01, variance_extra=0.1,
                                   tune_rate=2., n_iter_resample_v=None,
                                   second_order=True, use_vector_alpha=False)),
    }
    for name, (cls, kw) in samplers.items():
        s = cls(**kw)
        q = [q0.copy()]
        v0 = nz()
        if hasattr(s, "init_v"):
            s.init_v([v0])
        qs, ns, rs, mk, al = [], [], [], [], []
        for t in range(T):
            n, r = nz(), nz()
            if isinstance(s, (OS.SGHMC, OS.SGNHT)):
                q, info = s.step(q, model.grad, [r], [n])
            else:
                q, info = s.step(q, model.grad, [n])
            qs.append(q[0].copy()); ns.append(n); rs.append(r)
            if "mean_k" in info:
                mk.append(np.asarray(info["mean_k"][0], np.float32))
            if "alpha" in info:
                al.append(np.asarray(info["alpha"][0], np.float32))
        out[name + "_v0"] = v0
        out[name + "_q"] = np.stack(qs)
        out[name + "_noise"] = np.stack(ns)
        out[name + "_resample"] = np.stack(rs)
        if mk:
            out[name + "_mean_k"] = np.stack(mk)
        if al:
            out[name + "_alpha"] = np.stack(al)
    np.savez_compressed(os.path.join(HERE, "sgmcmc.npz"), **out)
    return samplers


SGMCMC_CONFIGS = None

if __name__ == "__main__":
    make_hmc_diag()
    make_hmc_dense()
    make_sgmcmc()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
