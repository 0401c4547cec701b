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


class PackedStats(object):
    """The per-iteration statistics message of a sharded HMC sampler and its collective schedule.

    Layout (float32): ``[sum(acc), n_chains, S1(D_0), S2(D_0), S1(D_1), ...]`` -- the acceptance
    part feeds ``reduce_mean(acceptance_rate)`` (hmc.py:377), the S1/S2 part the EWMV update
    (hmc.py:137-145), both global means over ALL chains.

    Schedule: the mass statistics that iteration t+1 consumes are those of the state AFTER
    iteration t's accept/reject, so they are computed right after the select and travel in the
    SAME all-reduce as iteration t's acceptance sum: ONE collective per iteration
    (``reduce_all``).  Only when no valid prefetch exists -- first iteration, the caller wrote
    the latent between two calls, adaptation was off in the previous call -- the mass part is
    reduced on its own at the start of the iteration (``reduce_mass``).  ``n_collectives``
    counts both, so tests and bench.py can assert the schedule."""

    def __init__(self, n_mass, device, group=None):
        self.buf = torch.zeros(2 + 2 * int(n_mass), dtype=torch.float32, device=device)
        self.acc = self.buf[:2]
        self.mass = self.buf[2:]
        self.group = group
        self.world = world(group)[0]
        self.n_collectives = 0
        self.mass_valid = False      # buf[2:] holds the GLOBAL sums for the next mass update

    def reduce_all(self, with_mass):
        """End of an iteration: acceptance sum (+ the prefetched mass statistics)."""
        if self.world > 1:
            all_reduce_sum(self.buf if with_mass else self.acc, self.group)
            self.n_collectives += 1
        self.mass_valid = bool(with_mass)

    def reduce_mass(self):
        """Start of an iteration without a valid prefetch."""
        if self.world > 1:
            all_reduce_sum(self.mass, self.group)
            self.n_collectives += 1
        self.mass_valid = True

    def reduce_acc(self):
        """Inside the step-size search loop (hmc.py:307-345): acceptance only."""
        if self.world > 1:
            all_reduce_sum(self.acc, self.group)
            self.n_collectives += 1


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


# ---- ELBO / IWAE: data-parallel over the batch axis -----------------------------
# The particle axis K stays local to a rank, so log_mean_exp needs no communication; the only
# exchange is the standard gradient all-reduce of the encoder/decoder parameters (5.4 MB at config
# 3) plus the scalar bound, packed into ONE flat buffer so there is one collective per step.
def all_reduce_weighted_mean_(mean, n_local, group=None):
    """In place: a LOCAL mean over ``n_local`` elements -> the mean over all ranks' elements
    (``sum_r mean_r * n_r / sum_r n_r``, one all-reduce of two floats).  No-op on one rank.
    Used by the scalar SGNHT thermostat, whose ``reduce_mean(v * v)`` (sgmcmc.py:494, 504)
    runs over ALL chains of the latent."""
    w, _ = world(group)
    if w == 1:
        return mean
    buf = torch.empty(2, dtype=torch.float32, device=mean.device)
    buf[0] = mean.reshape(-1)[0] * float(n_local)
    buf[1] = float(n_local)
    all_reduce_sum(buf, group)
    mean.reshape(-1)[0] = buf[0] / buf[1]
    return mean


def shard_batch(n_global, group=None):
    """Contiguous partition of the data/batch axis: (first index, local size)."""
    return shard_chains(n_global, group)


def all_reduce_mean_gradients(grads, extra_scalars=(), n_local=1, group=None):
    """Average per-datum-mean gradients over ranks with unequal shard sizes.

    ``grads``: gradients of the LOCAL mean cost (mean over this rank's ``n_local``
    data); ``extra_scalars``: local means to average the same way (the bound).
    Every tensor is weighted by ``n_local``, flattened into one buffer with the
    weight itself appended, summed with ONE all-reduce and divided by the
    global count -- exactly the gradient of the global mean.  Returns
    (list of averaged gradients, list of averaged scalars)."""
    w, _ = world(group)
    grads = list(grads)
    extra = [torch.as_tensor(s, dtype=torch.float32, device=grads[0].device).reshape(1)
             for s in extra_scalars]
    if w == 1:
        return grads, [e.reshape(()) for e in extra]
    flat = torch.cat([g.reshape(-1).to(torch.float32) * float(n_local) for g in grads] +
                     [e * float(n_local) for e in extra] +
                     [torch.full((1,), float(n_local), dtype=torch.float32,
                                 device=grads[0].device)])
    all_reduce_sum(flat, group)
    flat = flat / flat[-1]
    out, off = [], 0
    for g in grads:
        n = g.numel()
        out.append(flat[off:off + n].reshape(g.shape).to(g.dtype))
        off += n
    scal = [flat[off + i].reshape(()) for i in range(len(extra))]
    return out, scal
