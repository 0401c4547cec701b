"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous chain
sharding and the single all-reduce of the per-iteration statistics message
(section 8e): global mean acceptance and the EWMV mass update computed from
sharded statistics equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_all, acc_all, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from zhusuan_b200 import dist
    assert dist.world() == (world, rank)
    C, D = q_all.shape
    row0, n_local = dist.shard_chains(C)
    q = torch.tensor(q_all[row0:row0 + n_local])
    acc = torch.tensor(acc_all[row0:row0 + n_local])
    mean_old = torch.zeros(D)
    # what zsb_hmc_acc_sum_f32 / zsb_hmc_mass_stats_f32 write on each rank
    s1 = (q - mean_old).sum(0)
    s2 = ((q - mean_old) ** 2).sum(0)
    msg = dist.pack_stats(acc.sum(), n_local, s1, s2)
    dist.all_reduce_sum(msg)                     # the ONE collective / iteration
    abar = msg[0] / msg[1]
    n = msg[1]
    S1, S2 = msg[2:2 + D], msg[2 + D:2 + 2 * D]
    decay, tt = 0.99, 1.0
    w = (1 - decay) / (1 - decay ** tt)
    delta = S1 / n
    mean_new = mean_old + w * delta
    var_new = w * (S2 / n - w * delta * delta)   # (1-w)*0 + ...
    out[rank] = (float(abar), mean_new.numpy(), var_new.numpy(), row0, n_local)
    td.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_statistics_match_single_process_oracle():
    from oracle.hmc import ExponentialWeightedMovingVariance
    rng = np.random.RandomState(0)
    C, D = 37, 6                                  # ragged: 19 + 18 chains
    q_all = rng.standard_normal((C, D)).astype(np.float32)
    acc_all = rng.random_sample(C).astype(np.float32)
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, q_all, acc_all, out), nprocs=2, join=True)
    ew = ExponentialWeightedMovingVariance(0.99, [(1, D)], 1, np.float32)
    var = ew.update([q_all])[0].reshape(-1)
    assert out[0][3:] == (0, 19) and out[1][3:] == (19, 18)
    for r in (0, 1):
        abar, mean_new, var_new = out[r][:3]
        np.testing.assert_allclose(abar, acc_all.mean(), rtol=1e-6)
        np.testing.assert_allclose(mean_new, ew.mean[0].reshape(-1), rtol=1e-5,
                                   atol=1e-6)
        np.testing.assert_allclose(var_new, var, rtol=1e-4, atol=1e-6)


def _dp_worker(rank, world, port, x_all, w0, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from zhusuan_b200 import dist
    i0, n_local = dist.shard_batch(x_all.shape[0])
    x = torch.tensor(x_all[i0:i0 + n_local])
    w = torch.tensor(w0, requires_grad=True)
    b = torch.zeros((), requires_grad=True)
    # a stand-in objective with a per-datum mean cost (as mean(lower_bound.sgvb()), iwae.py:72-75)
    per_datum = torch.logsumexp(x @ w.t() + b, dim=1)
    cost = per_datum.mean()
    g = torch.autograd.grad(cost, [w, b])
    gs, (c,) = dist.all_reduce_mean_gradients(g, [cost.detach()], n_local=n_local)
    out[rank] = (gs[0].numpy(), gs[1].numpy(), float(c), i0, n_local)
    td.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_data_parallel_gradient_equals_global_mean_gradient():
    """Section 8e, ELBO/IWAE side: batch axis sharded 13 + 12, ONE all-reduce of the packed
    gradient + bound buffer reproduces the gradient of the global mean cost."""
    rng = np.random.RandomState(1)
    N, F, H = 25, 7, 5
    x_all = rng.standard_normal((N, F)).astype(np.float32)
    w0 = rng.standard_normal((H, F)).astype(np.float32)
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, x_all, w0, out), nprocs=2, join=True)
    w = torch.tensor(w0, requires_grad=True)
    b = torch.zeros((), requires_grad=True)
    cost = torch.logsumexp(torch.tensor(x_all) @ w.t() + b, dim=1).mean()
    gw, gb = torch.autograd.grad(cost, [w, b])
    assert (out[0][3], out[0][4], out[1][3], out[1][4]) == (0, 13, 13, 12)
    for r in (0, 1):
        np.testing.assert_allclose(out[r][0], gw.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out[r][1], gb.numpy(), rtol=1e-5, atol=1e-6)
        assert abs(out[r][2] - float(cost)) < 1e-5


def _packed_worker(rank, world, port, q_iters, acc_iters, out):
    """Drives zhusuan_b200.dist.PackedStats exactly as HMC._iterate_eager does (host restatement
    of what the acc_sum / mass_stats kernels write): ONE all-reduce per iteration carrying the
    acceptance sum of iteration t and the EWMV statistics iteration t+1 will consume."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from zhusuan_b200 import dist
    T_, C, D = q_iters.shape
    row0, n_local = dist.shard_chains(C)
    pk = dist.PackedStats(D, "cpu")
    mean, var = torch.zeros(D), torch.zeros(D)
    decay, tt = 0.99, 0.0
    res = []

    def local_mass_stats(q):
        pk.mass[:D] = (q - mean).sum(0)
        pk.mass[D:] = ((q - mean) ** 2).sum(0)

    for t in range(T_):
        q = torch.tensor(q_iters[t, row0:row0 + n_local])      # state at the START of iteration t
        if not pk.mass_valid:          # first iteration only: statistics + their own collective
            local_mass_stats(q)
            pk.reduce_mass()
        tt += 1.0
        n_glob = float(C)
        w = (1 - decay) / (1 - decay ** tt)
        delta = pk.mass[:D] / n_glob
        s2 = pk.mass[D:] / n_glob
        mean = mean + w * delta
        var = (1 - w) * var + w * (s2 - w * delta * delta)
        # ... trajectory + MH + select happen here; the post-select state is q_iters[t + 1]
        acc = torch.tensor(acc_iters[t, row0:row0 + n_local])
        pk.acc[0], pk.acc[1] = acc.sum(), float(n_local)
        if t + 1 < T_:
            local_mass_stats(torch.tensor(q_iters[t + 1, row0:row0 + n_local]))
        pk.reduce_all(with_mass=t + 1 < T_)                    # the ONE collective of iteration t
        res.append((float(pk.acc[0] / pk.acc[1]), mean.numpy().copy(), var.numpy().copy()))
    out[rank] = (res, pk.n_collectives)
    td.destroy_process_group()


@pytest.mark.timeout(120)
def test_packed_statistics_one_collective_per_iteration():
    """Section 8e: T iterations on 2 ranks cost T + 1 collectives (one extra for the very first
    mass update) and reproduce the single-process oracle's global mean acceptance and EWMV."""
    from oracle.hmc import ExponentialWeightedMovingVariance
    rng = np.random.RandomState(3)
    T_, C, D = 4, 23, 5
    q_iters = rng.standard_normal((T_, C, D)).astype(np.float32)
    acc_iters = rng.random_sample((T_, C)).astype(np.float32)
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_packed_worker, args=(2, port, q_iters, acc_iters, out), nprocs=2, join=True)
    ew = ExponentialWeightedMovingVariance(0.99, [(1, D)], 1, np.float32)
    for t in range(T_):
        var = ew.update([q_iters[t]])[0].reshape(-1)
        for r in (0, 1):
            abar, mean_t, var_t = out[r][0][t]
            np.testing.assert_allclose(abar, acc_iters[t].mean(), rtol=1e-6)
            np.testing.assert_allclose(mean_t, ew.mean[0].reshape(-1), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(var_t, var, rtol=1e-4, atol=1e-6)
    assert out[0][1] == out[1][1] == T_ + 1


def _wmean_worker(rank, world, port, v_all, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from zhusuan_b200 import dist
    c0, n_local = dist.shard_chains(v_all.shape[0])
    v = torch.tensor(v_all[c0:c0 + n_local])
    mk = (v * v).mean().reshape(1)                   # what zsb_sgmcmc_mean_sq_f32 leaves per rank
    dist.all_reduce_weighted_mean_(mk, v.numel())
    out[rank] = (float(mk[0]), c0, n_local)
    td.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_scalar_thermostat_sees_the_global_mean_kinetic_energy():
    """sgmcmc.py:494, 504: the scalar SGNHT thermostat is driven by reduce_mean(v * v) over ALL
    chains; with the chains sharded 7 + 6 the weighted all-reduce reproduces it on every rank."""
    rng = np.random.RandomState(3)
    v_all = (rng.standard_normal((13, 11)) * np.linspace(0.1, 3, 13)[:, None]).astype(np.float32)
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_wmean_worker, args=(2, port, v_all, out), nprocs=2, join=True)
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 7, 7, 6)
    for r in (0, 1):
        np.testing.assert_allclose(out[r][0], float((v_all.astype(np.float64) ** 2).mean()),
                                   rtol=1e-6)
    # one rank: the local mean is already the global one
    from zhusuan_b200 import dist
    m = torch.tensor([0.25])
    assert float(dist.all_reduce_weighted_mean_(m, 10)[0]) == 0.25
