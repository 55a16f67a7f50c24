"""Padded-adjacency construction - the input contract of the sampler
(reference graphsage/minibatch.py:227-259, NodeMinibatchIterator.construct_adj / construct_test_adj).

Host-side, start-up-time code (the reference builds it once in numpy too).  networkx-free:
the graph is given as CSR over node indices 0..N-1.

  adj[N+1, max_degree] int32, filled with N; row N is the dummy row; a node with no usable
  neighbours keeps an all-N row; deg > max_degree -> subsample without replacement,
  deg < max_degree -> resample with replacement.
"""
import numpy as np


def _pad_row(nb, max_degree, rng):
    if len(nb) > max_degree:
        return rng.choice(nb, max_degree, replace=False)
    if len(nb) < max_degree:
        return rng.choice(nb, max_degree, replace=True)
    return nb


def construct_adj(indptr, indices, max_degree, skip_node=None, edge_removed=None, node_order=None, rng=None):
    """Train-time table (reference minibatch.py:227-245).

    skip_node[N] bool   : val/test nodes - their rows stay all-N (:232-233)
    edge_removed[nnz] bool : per CSR entry, the edge's `train_removed` flag (:234-236)
    node_order          : iteration order of nodes (decides the RNG call order); default 0..N-1
    rng                 : numpy RandomState; default RandomState(123) (reference seeds 123, minibatch.py:6)
    Returns (adj int32 [N+1, max_degree], deg float64 [N]).
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices)
    n = len(indptr) - 1
    rng = np.random.RandomState(123) if rng is None else rng
    adj = np.full((n + 1, max_degree), n, dtype=np.int32)
    deg = np.zeros((n,), dtype=np.float64)
    order = range(n) if node_order is None else node_order
    for u in order:
        if skip_node is not None and skip_node[u]:
            continue
        lo, hi = indptr[u], indptr[u + 1]
        nb = indices[lo:hi]
        if edge_removed is not None:
            nb = nb[~np.asarray(edge_removed[lo:hi], dtype=bool)]
        deg[u] = len(nb)
        if len(nb) == 0:
            continue
        adj[u, :] = _pad_row(nb, max_degree, rng)
    return adj, deg


def construct_test_adj(indptr, indices, max_degree, node_order=None, rng=None):
    """Test-time table over ALL edges (reference minibatch.py:247-259)."""
    adj, _ = construct_adj(indptr, indices, max_degree, None, None, node_order, rng)
    return adj


def padded_from_csr_fast(indptr, indices, max_degree, seed=123):
    """Vectorised construction with the same distribution (not the same RNG stream) for large
    synthetic graphs: rows with deg >= max_degree keep a uniform random subset without replacement,
    rows with 0 < deg < max_degree are filled by uniform draws with replacement."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int32)
    n = len(indptr) - 1
    rs = np.random.RandomState(seed)
    deg = np.diff(indptr)
    adj = np.full((n + 1, max_degree), n, dtype=np.int32)
    small = np.nonzero((deg > 0) & (deg < max_degree))[0]
    if len(small):
        pos = (rs.random_sample((len(small), max_degree)) * deg[small, None]).astype(np.int64)
        adj[small] = indices[indptr[small, None] + pos]
    big = np.nonzero(deg >= max_degree)[0]
    if len(big):
        # random keys per entry, take the max_degree smallest keys of each row
        rows = np.repeat(big, deg[big])
        ent = np.concatenate([np.arange(indptr[u], indptr[u + 1]) for u in big]) if len(big) < 4096 else \
            (np.arange(deg[big].sum()) - np.repeat(np.cumsum(deg[big]) - deg[big], deg[big]) + np.repeat(indptr[big], deg[big]))
        keys = rs.random_sample(len(ent))
        order = np.lexsort((keys, rows))
        ent, rows = ent[order], rows[order]
        start = np.cumsum(deg[big]) - deg[big]
        rank = np.arange(len(ent)) - np.repeat(start, deg[big])
        keep = rank < max_degree
        adj[rows[keep], rank[keep]] = indices[ent[keep]]
    return adj, deg.astype(np.float64)
