"""Oracle for the device-side R-MAT CSR generator (graphsage_b200/csrc/rmat.cu, gs_rmat_degrees / gs_rmat_fill).

No reference counterpart: the reference loads graphs from disk (graphsage/utils.py:19-75); BASELINE.json configs[4] asks
for a 100 M-node R-MAT graph (a, b, c, d = 0.57, 0.19, 0.19, 0.05), which has to be synthesised.  The construction -
row first (degree from the row marginal), then the column's bits given the row's bits - and the Philox contract are
documented at the top of csrc/rmat.cu; this file restates them in numpy, vectorised over all entries, for sizes a CPU
finishes in seconds.

Test infrastructure - not imported by the product.
"""
import math

import numpy as np

from .philox import philox4x32_10, split64

TAG_DEG = 0x08000000
TAG_FILL = 0x08000001


def scramble_constants(n):
    """(mul, mul_inv, add) of the id bijection y = (x * mul + add) mod n used by the product (graphsage_b200/synthetic.py)."""
    mul = 0x9E3779B1 % n
    if mul < 2:
        mul = 1
    while math.gcd(mul, n) != 1:
        mul += 1
    return mul, pow(mul, -1, n), 0x7F4A7C15 % n


def _popcount(x):
    x = np.asarray(x, dtype=np.uint64)
    c = np.zeros(x.shape, dtype=np.int64)
    for i in range(40):
        c += ((x >> np.uint64(i)) & np.uint64(1)).astype(np.int64)
    return c


def _prow(scale, a, b, c, d):
    out = []
    for z in range(scale + 1):
        v = 1.0
        for _ in range(scale - z):
            v = v * (a + b)
        for _ in range(z):
            v = v * (c + d)
        out.append(v)
    return np.array(out, dtype=np.float64)


def rmat_csr(scale, n_nodes, edge_factor=20.0, a=0.57, b=0.19, c=0.19, d=0.05, seed=123):
    """Returns (indptr int64 [n+1], indices int32) exactly as gs_rmat_degrees + prefix sum + gs_rmat_fill produce them."""
    n = int(n_nodes)
    mul, mul_inv, add = scramble_constants(n)
    prow = _prow(scale, a, b, c, d)
    space = 1 << scale
    key = np.array(split64(seed), dtype=np.uint32)
    y = np.arange(n, dtype=np.int64)
    r = ((y - add) % n) * mul_inv % n
    r2 = r + n
    has2 = r2 < space
    p1 = prow[_popcount(r)]
    p2 = np.where(has2, prow[np.minimum(_popcount(r2), scale)], 0.0)
    lam = (edge_factor * float(n)) * np.where(has2, p1 + p2, p1)
    fl = np.floor(lam)
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = y.astype(np.uint32)
    ctr[:, 3] = TAG_DEG
    u = (philox4x32_10(ctr, key)[:, 0].astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    deg = np.minimum(fl + (u < (lam - fl)), 2147483647.0).astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    m = int(indptr[-1])
    rows = np.repeat(y, deg)
    j = np.arange(m, dtype=np.int64) - np.repeat(indptr[:-1], deg)
    hw = np.zeros((m, 32), dtype=np.uint32)
    for blk in range(4):
        ctr = np.zeros((m, 4), dtype=np.uint32)
        ctr[:, 0] = j.astype(np.uint32)
        ctr[:, 1] = blk
        ctr[:, 2] = rows.astype(np.uint32)
        ctr[:, 3] = TAG_FILL
        w = philox4x32_10(ctr, key)
        for q in range(4):
            hw[:, blk * 8 + 2 * q] = w[:, q] & 0xFFFF
            hw[:, blk * 8 + 2 * q + 1] = w[:, q] >> 16
    pick_thr = np.where(has2, (65536.0 * (p2 / np.where(has2, p1 + p2, 1.0))).astype(np.uint32), 0).astype(np.uint32)
    x = np.where(hw[:, 0] < pick_thr[rows], r2[rows], r[rows])
    thr0, thr1 = int(65536.0 * (b / (a + b))), int(65536.0 * (d / (c + d)))
    col = np.zeros(m, dtype=np.int64)
    for lvl in range(scale):
        rb = (x >> (scale - 1 - lvl)) & 1
        cb = (hw[:, 1 + lvl] < np.where(rb == 1, thr1, thr0)).astype(np.int64)
        col = (col << 1) | cb
    col %= n
    cy = (col * mul + add) % n
    cy = np.where(cy == rows, (cy + 1) % n, cy)
    return indptr, cy.astype(np.int32)
