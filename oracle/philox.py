"""Philox4x32-10 counter-based RNG in NumPy (TEST ORACLE ONLY).

The reference draws noise with TensorFlow's stateful Philox ops
(``tf.random_normal`` hmc.py:22, sgmcmc.py:196; ``tf.random_uniform``
hmc.py:485) whose streams cannot be reproduced without TF, so parity runs
inject noise.  The product's in-kernel generator is the published
Philox4x32-10 (Salmon et al., SC'11, Random123) keyed by
(seed; stream, iteration, global chain, element block); this file restates
that algorithm so the in-kernel draws can be checked bit-for-bit, and is
itself pinned to the Random123 known-answer vectors in
tests/test_oracle_pins.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32[..., 4], key: uint32[..., 2] -> uint32[..., 4]."""
    ctr = np.asarray(ctr, np.uint32)
    key = np.asarray(key, np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0 = np.broadcast_to(key[..., 0], c0.shape).astype(np.uint32)
    k1 = np.broadcast_to(key[..., 1], c0.shape).astype(np.uint32)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u32_to_uniform(x):
    """uint32 -> float32 in [0, 1): (x >> 8) * 2^-24 (24 mantissa bits)."""
    return ((np.asarray(x, np.uint32) >> np.uint32(8)).astype(np.float32)
            * np.float32(1.0 / 16777216.0))


def u32_to_uniform_open(x):
    """uint32 -> float32 in (0, 1]: ((x >> 8) + 1) * 2^-24 (safe for log)."""
    return (((np.asarray(x, np.uint32) >> np.uint32(8)).astype(np.float32)
             + np.float32(1.0)) * np.float32(1.0 / 16777216.0))


def box_muller(u_a, u_b):
    """Two uint32 words -> two float32 standard normals.

    r = sqrt(-2 ln u1), u1 in (0,1];  theta = 2 pi u2, u2 in [0,1).
    """
    u1 = u32_to_uniform_open(u_a).astype(np.float64)
    u2 = u32_to_uniform(u_b).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * np.pi * u2
    return ((r * np.cos(th)).astype(np.float32),
            (r * np.sin(th)).astype(np.float32))


def counter(stream, iteration, row, block):
    """The product's counter layout: (block, row, iteration, stream)."""
    row, block = np.broadcast_arrays(np.asarray(row, np.uint32),
                                     np.asarray(block, np.uint32))
    c = np.empty(row.shape + (4,), np.uint32)
    c[..., 0] = block
    c[..., 1] = row
    c[..., 2] = np.uint32(iteration)
    c[..., 3] = np.uint32(stream)
    return c


def normal_matrix(seed, stream, iteration, row0, n_rows, n_cols):
    """Standard normals for rows [row0, row0+n_rows) x n_cols, exactly as the
    kernels draw them: element (r, c) uses Philox block c // 4 of global row
    r, word c % 4; words (0,1) and (2,3) feed one Box-Muller pair each."""
    nblk = (n_cols + 3) // 4
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row0)).astype(
        np.uint32)[:, None]
    blks = np.arange(nblk, dtype=np.uint32)[None, :]
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    w = philox4x32_10(counter(stream, iteration, rows, blks), key)
    z0, z1 = box_muller(w[..., 0], w[..., 1])
    z2, z3 = box_muller(w[..., 2], w[..., 3])
    out = np.stack([z0, z1, z2, z3], axis=-1).reshape(n_rows, nblk * 4)
    return out[:, :n_cols]


def uniform_vector(seed, stream, iteration, row0, n_rows):
    """One uniform [0,1) per global row: word 0 of Philox block 0."""
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row0)).astype(
        np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    w = philox4x32_10(counter(stream, iteration, rows, np.uint32(0)), key)
    return u32_to_uniform(w[..., 0])
