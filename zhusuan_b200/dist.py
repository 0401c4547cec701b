"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).

The samplers shard the chain axis across ranks with NO data-path collective;
the only exchange is one all-reduce(sum) of a tiny statistics buffer per HMC
iteration -- [sum(acc), n_chains] (8 B), plus [S1(D), S2(D)] while the mass is
adapting (hmc.py:377 ``reduce_mean(acceptance_rate)`` and hmc.py:137-145 are
global means over ALL chains).
"""
import torch


def world(group=None):
    """(world_size, rank) of ``group`` (1, 0 when torch.distributed is not
    initialised)."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        return td.get_world_size(group), td.get_rank(group)
    return 1, 0


def all_reduce_sum(t, group=None):
    import torch.distributed as td
    td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t


def shard_chains(n_chains_global, group=None):
    """Contiguous partition of the flattened chain axis: returns
    (row0, n_local) for this rank; remainders go to the lowest ranks."""
    w, r = world(group)
    base, rem = divmod(int(n_chains_global), w)
    n_local = base + (1 if r < rem else 0)
    row0 = r * base + min(r, rem)
    return row0, n_local


def pack_stats(acc_sum, n_local, s1=None, s2=None):
    """The per-iteration statistics message (host-side helper used by the
    gloo tests to restate what the kernels write into the stats buffers)."""
    parts = [torch.as_tensor([float(acc_sum), float(n_local)],
                             dtype=torch.float32)]
    if s1 is not None:
        parts += [s1.reshape(-1).float(), s2.reshape(-1).float()]
    return torch.cat(parts)
