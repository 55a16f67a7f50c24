"""Synthetic graphs shaped like the reference's datasets (no network, no real data).

reddit_like(): N=232,965 nodes, 41 communities, heavy-tailed degrees (mean ~ 50 undirected
neighbours before symmetrisation => ~100 after, like Reddit's 11.6 M undirected edges), F=602
N(0,1) features.  Used by bench.py and the full-size GPU tests (BASELINE.json configs[1..3]).
"""
import numpy as np


def community_graph_csr(n, n_comm=41, mean_deg=50, p_in=0.8, seed=123, max_deg_cap=2000):
    """Undirected community-structured graph with a Pareto degree tail, returned as CSR
    (indptr int64 [n+1], indices int32) over nodes 0..n-1 with sorted, de-duplicated rows."""
    rs = np.random.RandomState(seed)
    comm = rs.randint(0, n_comm, size=n).astype(np.int32)
    order = np.argsort(comm, kind="stable")
    comm_start = np.searchsorted(comm[order], np.arange(n_comm))
    comm_size = np.diff(np.append(comm_start, n))
    # out-degrees: Pareto(alpha=2.2) scaled to the target mean, capped
    raw = (rs.pareto(2.2, size=n) + 1.0)
    d = np.minimum(np.maximum((raw * (mean_deg / raw.mean())).astype(np.int64), 1), max_deg_cap)
    src = np.repeat(np.arange(n, dtype=np.int64), d)
    m = len(src)
    inside = rs.random_sample(m) < p_in
    c = comm[src]
    dst_in = order[comm_start[c] + (rs.random_sample(m) * comm_size[c]).astype(np.int64)]
    dst_out = rs.randint(0, n, size=m)
    dst = np.where(inside, dst_in, dst_out).astype(np.int64)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    a = np.concatenate([src, dst])
    b = np.concatenate([dst, src])
    key = a * n + b
    key.sort()
    key = key[np.concatenate([[True], key[1:] != key[:-1]])]
    a, b = key // n, (key % n).astype(np.int32)
    indptr = np.concatenate([[0], np.cumsum(np.bincount(a, minlength=n))]).astype(np.int64)
    return indptr, b, comm


def reddit_like(n=232965, f=602, max_degree=128, seed=123, with_features=True, mean_deg=50, locality=True):
    """Returns dict(indptr, indices, adj[n+1, max_degree] int32, deg, comm, features[n+1, f] fp32 with zero last row)."""
    from .minibatch import padded_from_csr_fast
    from .parallel import locality_order, relabel_graph
    indptr, indices, comm = community_graph_csr(n, mean_deg=mean_deg, seed=seed)
    if locality:
        # rename nodes so communities are contiguous id ranges (what a locality-aware node partition wants)
        order, inv = locality_order(comm)
        indptr, indices = relabel_graph(indptr, indices, order, inv)
        comm = comm[order]
    adj, deg = padded_from_csr_fast(indptr, indices, max_degree, seed=seed)
    out = dict(indptr=indptr, indices=indices, adj=adj, deg=deg, comm=comm, n=n, f=f, max_degree=max_degree)
    if with_features:
        rs = np.random.RandomState(seed + 1)
        feats = np.zeros((n + 1, f), dtype=np.float32)
        feats[:n] = rs.standard_normal((n, f)).astype(np.float32)
        out["features"] = feats
    return out


def rmat_csr(scale, edge_factor=20, a=0.57, b=0.19, c=0.19, d=0.05, n_nodes=None, seed=123, undirected=False,
             chunk=1 << 24):
    """R-MAT graph (a, b, c, d = 0.57, 0.19, 0.19, 0.05; scale 27 trimmed to 10^8 nodes,
    about 20 directed entries per node) as CSR over nodes 0..n-1 with sorted, de-duplicated rows and no self loops.

    Each of the `edge_factor * n` directed edges picks one quadrant per bit level with probabilities (a, b, c, d); ids
    that fall beyond `n_nodes` (when 2^scale is trimmed) are redrawn by folding (id mod n_nodes).  Generated in chunks
    so the peak host memory stays bounded; deterministic for a given (scale, edge_factor, seed)."""
    assert abs(a + b + c + d - 1.0) < 1e-9
    n = int(n_nodes) if n_nodes is not None else 1 << scale
    m = int(edge_factor) * n
    rs = np.random.RandomState(seed)
    keys = []
    done = 0
    while done < m:
        cnt = min(chunk, m - done)
        src = np.zeros(cnt, dtype=np.int64)
        dst = np.zeros(cnt, dtype=np.int64)
        for _ in range(scale):
            r = rs.random_sample(cnt)
            right = (r >= a) & (r < a + b) | (r >= a + b + c)          # quadrants b and d: destination bit set
            down = r >= a + b                                            # quadrants c and d: source bit set
            src = (src << 1) | down
            dst = (dst << 1) | right
        if n != (1 << scale):
            src %= n
            dst %= n
        keep = src != dst
        src, dst = src[keep], dst[keep]
        if undirected:
            src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
        keys.append(src * n + dst)
        done += cnt
    key = np.unique(np.concatenate(keys))
    rows, cols = key // n, (key % n).astype(np.int32)
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64)
    return indptr, cols


def rmat_scramble_constants(n):
    """(mul, mul_inv, add) of the node-id bijection y = (x * mul + add) mod n applied by the device generator (R-MAT puts
    its hubs on ids with few set bits; contiguous-range partitions would otherwise give every hub to GPU 0)."""
    import math
    mul = 0x9E3779B1 % n
    if mul < 2:
        mul = 1
    while math.gcd(mul, n) != 1:
        mul += 1
    return mul, pow(mul, -1, n), 0x7F4A7C15 % n


def rmat_csr_device(scale, n_nodes=None, edge_factor=20.0, a=0.57, b=0.19, c=0.19, d=0.05, seed=123, device="cuda",
                    long_threshold=4096):
    """R-MAT graph written directly as CSR on the device (gs_rmat_degrees -> prefix sum -> gs_rmat_fill; contract in
    csrc/rmat.cu / oracle/rmat.py).  Returns (indptr int64 [n+1], indices int32 [m]) CUDA tensors.  BASELINE configs[4]:
    scale=27, n_nodes=100_000_000 (8 GB of indices)."""
    import torch
    from . import _lib
    from ._lib import check, lib, ptr, stream_ptr
    n = int(n_nodes) if n_nodes else 1 << scale
    mul, mul_inv, add = rmat_scramble_constants(n)
    dev = torch.device(device)
    deg = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().gs_rmat_degrees(scale, n, float(edge_factor), a, b, c, d, seed & (2**64 - 1), mul, mul_inv, add, ptr(deg),
                                    stream_ptr()))
        indptr = torch.zeros((n + 1,), dtype=torch.int64, device=dev)
        torch.cumsum(deg, dim=0, out=indptr[1:])
        m = int(indptr[-1].item())
        indices = torch.empty((m,), dtype=torch.int32, device=dev)
        threshold = int(long_threshold)                    # rows beyond this are filled by a whole grid, not one warp
        long_rows = torch.nonzero(deg > threshold).reshape(-1)
        while long_rows.numel() > 65535:
            threshold *= 2
            long_rows = torch.nonzero(deg > threshold).reshape(-1)
        check(lib().gs_rmat_fill(scale, n, a, b, c, d, seed & (2**64 - 1), mul, mul_inv, add, ptr(indptr), ptr(indices),
                                 ptr(long_rows) if long_rows.numel() else 0, long_rows.numel(), threshold, stream_ptr()))
    return indptr, indices
