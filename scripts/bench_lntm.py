#!/usr/bin/env python
"""Config 5 (BASELINE.json): Logistic-Normal Topic Model HMC inner loop, 10 000-document synthetic
corpus, 128 topics, 1024 chains, V = 8192 (examples/topic_models/lntm_mcem.py:33-48, 97-105).
Times the fused sparsity-aware log-joint kernel (value + gradient) and one HMC iteration on it;
unit: chain-document gradient evaluations / s and leapfrog-steps*chains/s (a "chain" here is one
(chain, document) pair: the reference's chain axes are [n_chains, n_docs])."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zhusuan_b200 as zs  # noqa: E402


def main():
    dev = torch.device("cuda")
    C = int(os.environ.get("LNTM_CHAINS", 1024))
    Dn, K, V, words = 10000, 128, 8192, 150
    rng = np.random.Generator(np.random.PCG64(3))
    # synthetic bag of words: ~150 distinct words per document, Zipf-ish counts
    rows = np.repeat(np.arange(Dn), words)
    cols = rng.integers(0, V, Dn * words)
    x = torch.zeros(Dn, V, device=dev)
    x.index_put_((torch.tensor(rows, device=dev), torch.tensor(cols, device=dev)),
                 torch.tensor(1.0 + rng.poisson(0.5, Dn * words), dtype=torch.float32, device=dev),
                 accumulate=True)
    beta = torch.randn(K, V, device=dev)
    lj = zs.fused.LNTMLogJoint(x, beta, torch.zeros(K, device=dev), torch.zeros(K, device=dev))
    nnz = int(lj.doc_ptr[-1])
    eta = 0.1 * torch.randn(C, Dn, K, device=dev)
    for _ in range(2):
        lj.grad([eta])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 5
    e0.record()
    for _ in range(n):
        g = lj.grad([eta])[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 4.0 * K * nnz * C            # dot + axpy per (chain, word occurrence)
    out = {"workload": "LNTM E-step log-joint gradient, %d chains x %d docs x %d topics, V=%d, "
                       "nnz=%d" % (C, Dn, K, V, nnz),
           "grad_ms": ms, "chain_doc_grads_per_s": C * Dn / (ms * 1e-3),
           "fp32_tflops": flops / (ms * 1e-3) / 1e12,
           "dense_equivalent_tflops": 4.0 * K * V * C * Dn / (ms * 1e-3) / 1e12,
           "eta_bytes_GBps": 2 * eta.numel() * 4 / (ms * 1e-3) / 1e9}
    L = 5
    h = zs.HMC(step_size=1e-3, n_leapfrogs=L, adapt_step_size=True, target_acceptance_rate=0.6,
               seed=3)
    op, info = h.sample(lj, {}, {"eta": eta})
    for _ in range(2):
        op()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        op()
    e1.record()
    op.synchronize()
    ms_it = e0.elapsed_time(e1) / 3
    out["hmc_ms_per_iteration_L%d" % L] = ms_it
    out["leapfrog_steps_chain_docs_per_s"] = C * Dn * L / (ms_it * 1e-3)
    out["acceptance_mean"] = float(info.acceptance_rate.mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
