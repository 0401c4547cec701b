"""CPU baseline for bench.py: the reference's HMC iteration restated op-by-op
in torch-CPU (TEST/BENCH INFRASTRUCTURE ONLY -- never a fallback).

TensorFlow is not installable here, so the reference
itself cannot be timed.  This restates the graph ``HMC.sample`` builds
(zhusuan/hmc.py:382-522) with the same un-fused op sequence TF-CPU would
execute: fresh ``randn`` momentum, L+1 leapfrog passes each with ONE reverse-
mode gradient of the user log-joint (hmc.py:430-432), two extra forward
evaluations for the MH test (hmc.py:47-50), ``where`` select -- using every
host thread torch will give it.  ``kind`` is therefore "port".
"""
import time

import numpy as np
import torch


def dense_gaussian_log_joint(P, const):
    def log_joint(x):
        return -0.5 * ((x @ P) * x).sum(-1) + const
    return log_joint


def hmc_iteration_torch_cpu(log_joint, q, n_leapfrogs, step_size, mass):
    """One iteration, adaptation off (the per-iteration adaptation cost is two
    reductions over q, negligible beside L+1 gradients)."""
    def grad(x):
        x = x.detach().requires_grad_(True)
        lp = log_joint(x)
        return torch.autograd.grad(lp.sum(), x)[0]

    p = torch.randn_like(q) * torch.sqrt(mass)              # hmc.py:21-23
    cq, cp = q, p
    L = n_leapfrogs
    for i in range(L + 1):                                  # hmc.py:352-364
        s1 = step_size if i > 0 else 0.0
        s2 = step_size if 0 < i < L else step_size / 2
        cq = cq + s1 * (cp / mass)
        cp = cp + s2 * grad(cq)
    with torch.no_grad():                                   # hmc.py:46-61
        lp0, lp1 = log_joint(q), log_joint(cq)
        h0 = -lp0 + 0.5 * (p * p / mass).sum(-1)
        h1 = -lp1 + 0.5 * (cp * cp / mass).sum(-1)
        acc = torch.exp(torch.minimum(h0 - h1, torch.zeros_like(h0)))
        acc = torch.where(torch.isfinite(acc) & torch.isfinite(lp1), acc,
                          torch.zeros_like(acc))
        u = torch.rand_like(acc)
        accept = (u < acc).unsqueeze(-1)
        new_q = torch.where(accept, cq, q)                  # hmc.py:488-497
    return new_q, acc


def time_dense_hmc(D, chains, n_leapfrogs, n_iters, warmup=1, seed=2,
                   step_size=0.05, P=None):
    """Returns dict(value=leapfrog-steps*chains/s, seconds, cores, sample)."""
    from .models import make_dense_gaussian_problem
    torch.manual_seed(seed)
    if P is None:
        P64, const = make_dense_gaussian_problem(D, seed=seed)
    else:
        P64, const = P
    Pt = torch.tensor(P64, dtype=torch.float32)
    lj = dense_gaussian_log_joint(Pt, float(const))
    q = torch.randn(chains, D)
    mass = torch.ones(D)
    for _ in range(warmup):
        q, _ = hmc_iteration_torch_cpu(lj, q, n_leapfrogs, step_size, mass)
    t0 = time.perf_counter()
    for _ in range(n_iters):
        q, acc = hmc_iteration_torch_cpu(lj, q, n_leapfrogs, step_size, mass)
    dt = time.perf_counter() - t0
    units = chains * n_leapfrogs * n_iters
    return {"value": units / dt, "seconds": dt,
            "cores": torch.get_num_threads(),
            "sample": "%d chains x %d-d, L=%d, %d iteration(s), torch-CPU "
                      "restatement of hmc.py:382-522 (unfused, autograd "
                      "gradient per leapfrog pass)" % (chains, D, n_leapfrogs,
                                                       n_iters),
            "ms_per_iter": 1e3 * dt / n_iters}
