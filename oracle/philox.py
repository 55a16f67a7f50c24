"""Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as
1, 2, 3"; Random123 v1.09 ``philox4x32_R(10, ctr, key)``) in vectorised numpy.

This is the RNG contract shared by the oracle and the CUDA sampler kernels
(graphsage_b200/csrc/philox.cuh).  It replaces TensorFlow's
``tf.random_shuffle`` stream (reference graphsage/neigh_samplers.py:27), whose
exact values depend on TF graph-construction state and cannot be reproduced
without TensorFlow: "parity unpinned" for the stream,
pinned for everything computed from it.

Test infrastructure - not imported by the product.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 (broadcastable) -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    shape = np.broadcast_shapes(ctr.shape[:-1], key.shape[:-1])
    c = [np.broadcast_to(ctr[..., i], shape).astype(np.uint64) for i in range(4)]
    k0 = np.broadcast_to(key[..., 0], shape).astype(np.uint32)
    k1 = np.broadcast_to(key[..., 1], shape).astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[0]
            p1 = _M1 * c[2]
            hi0, lo0 = p0 >> _S32, p0 & _MASK
            hi1, lo1 = p1 >> _S32, p1 & _MASK
            c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1,
                 hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def mulhi32(a, b):
    """floor(a*b / 2**32) for uint32 a, b: maps a uniform u32 onto [0, b)."""
    return ((np.asarray(a, dtype=np.uint64) * np.asarray(b, dtype=np.uint64)) >> _S32).astype(np.uint32)


def split64(x):
    x = int(x) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(x & 0xFFFFFFFF), np.uint32(x >> 32)
