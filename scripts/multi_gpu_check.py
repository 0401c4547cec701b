#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/multi_gpu_check.py

Sharded-vs-single parity: N ranks each hold a contiguous shard of the chains
and run adaptive HMC with in-kernel Philox noise (keyed by GLOBAL chain index);
rank 0 also runs all chains on one GPU.  The only cross-rank exchange is the
per-iteration statistics all-reduce, so samples must agree to fp32 rounding of
that sum (step sizes within 1e-5 relative, samples within 1e-3 after 12
adaptive iterations) and acceptance decisions must match almost everywhere.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zhusuan_b200 as zs                               # noqa: E402
from bench import make_dense_gaussian_problem          # noqa: E402  (synthetic target)


def run(model_fn, q, n_iters, group, chain_offset, dense, graph=False):
    kw = dict(step_size=0.05 if dense else 1e-3, n_leapfrogs=6,
              adapt_step_size=True, adapt_mass=True, mass_collect_iters=4,
              seed=99, process_group=group, chain_offset=chain_offset,
              use_cuda_graph=graph)
    h = zs.HMC(**kw)
    op, info = h.sample(model_fn, {}, {"x": q})
    for i in range(n_iters):
        op(adapt_step_size=True, adapt_mass=True)
    op.synchronize()
    c_adapt = h._pk.n_collectives
    for i in range(3):                      # sampling phase: nothing adapts, nothing to exchange
        op(adapt_step_size=False, adapt_mass=False)
    op.synchronize()
    info.n_collectives = c_adapt
    info.n_collectives_sampling = h._pk.n_collectives - c_adapt
    info.n_search = h.n_search_iters
    return info


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    ok = True
    for case in ("diag", "dense64", "dense1024", "dense1024-graph"):
        dense = case != "diag"
        graph = case.endswith("graph")
        D, C = {"diag": (100, 64 * world), "dense64": (64, 96 * world)}.get(
            case, (1024, 320 * world))
        g = torch.Generator(device="cpu"); g.manual_seed(5)
        q_all = torch.randn(C, D, generator=g) * 0.5
        if dense:
            P, _ = make_dense_gaussian_problem(D, seed=3)
            model = zs.fused.GaussianLogJoint(P, device=dev)
        else:
            std = torch.tensor(1.0 / (1.0 + np.arange(D)), dtype=torch.float32, device=dev)

            @zs.meta_bayesian_net()
            def gaussian():
                bn = zs.BayesianNet()
                bn.normal('x', torch.zeros(D, device=dev), std=std, group_ndims=1)
                return bn
            model = gaussian()
        n_local = C // world
        q = q_all[rank * n_local:(rank + 1) * n_local].to(dev).contiguous()
        info = run(model, q, 12, None, None, dense, graph)   # sharded (default group)
        gathered = [torch.empty_like(q) for _ in range(world)]
        td.all_gather(gathered, q)
        accs = [torch.empty_like(info.acceptance_rate) for _ in range(world)]
        td.all_gather(accs, info.acceptance_rate.contiguous())
        if rank == 0:
            # single-GPU run of ALL chains: a 1-rank group disables the all-reduce
            solo = td.new_group([0])
        else:
            solo = td.new_group([0])
        if rank == 0:
            q1 = q_all.to(dev).contiguous()
            info1 = run(model, q1, 12, solo, 0, dense)
            qs = torch.cat(gathered)
            d = (qs - q1).abs().max().item()
            same_rows = ((qs - q1).abs().amax(1) < 1e-3).float().mean().item()
            ss = abs(float(info.updated_step_size) - float(info1.updated_step_size)) / float(info1.updated_step_size)
            # ONE packed all-reduce per iteration (+1 for the very first mass update, + the
            # acceptance-only reductions inside the two step-size searches)
            want = 12 + 1 + info.n_search
            print("%s: max|dq| %.3e  rows equal %.4f  step-size rel diff %.2e  acc mean %.4f vs "
                  "%.4f  collectives %d (expected %d)"
                  % (case, d, same_rows, ss, torch.cat(accs).mean().item(),
                     info1.acceptance_rate.mean().item(), info.n_collectives, want))
            ok &= same_rows > 0.98 and ss < 1e-4 and info.n_collectives == want
            ok &= info.n_collectives_sampling == 0      # 3 non-adapting iterations: no all-reduce
            print("   + 3 non-adapting iterations: %d collectives (expected 0)"
                  % info.n_collectives_sampling)
        td.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    td.destroy_process_group()


if __name__ == "__main__":
    main()
