"""Oracle for the padded-adjacency contract the sampler consumes
(reference graphsage/minibatch.py:227-259, NodeMinibatchIterator.construct_adj /
construct_test_adj).

Graph form used here (networkx-free): `nodes` = node ids in iteration order,
`neighbors[node]` = list of neighbour ids in iteration order, `is_val_or_test[node]`
bool, `train_removed[(u, v)]` bool per directed pair, `id2idx[node]` = row.
`rng` is a numpy legacy RandomState; the reference seeds the global one with 123
(graphsage/minibatch.py:6) and calls np.random.choice in node-iteration order.

Test infrastructure - not imported by the product.
"""
import numpy as np


def _pad(neighbors, max_degree, rng):
    if len(neighbors) > max_degree:                                   # :240-241
        return rng.choice(neighbors, max_degree, replace=False)
    if len(neighbors) < max_degree:                                   # :242-243
        return rng.choice(neighbors, max_degree, replace=True)
    return neighbors


def construct_adj(nodes, neighbors, id2idx, is_val_or_test, train_removed, max_degree, rng):
    n = len(id2idx)
    adj = n * np.ones((n + 1, max_degree))                            # :228  (dummy row n, fill value n)
    deg = np.zeros((n,))                                              # :229
    for nodeid in nodes:
        if is_val_or_test[nodeid]:                                    # :232-233
            continue
        nb = np.array([id2idx[v] for v in neighbors[nodeid]
                       if not train_removed.get((nodeid, v), False)])  # :234-236
        deg[id2idx[nodeid]] = len(nb)                                 # :237
        if len(nb) == 0:
            continue
        adj[id2idx[nodeid], :] = _pad(nb, max_degree, rng)            # :240-244
    return adj.astype(np.int32), deg


def construct_test_adj(nodes, neighbors, id2idx, max_degree, rng):
    n = len(id2idx)
    adj = n * np.ones((n + 1, max_degree))                            # :248
    for nodeid in nodes:
        nb = np.array([id2idx[v] for v in neighbors[nodeid]])         # :250-251
        if len(nb) == 0:
            continue
        adj[id2idx[nodeid], :] = _pad(nb, max_degree, rng)
    return adj.astype(np.int32)


def build_padded_adj(indptr, indices, max_degree, seed, counter, skip=None):
    """Device-builder contract (graphsage_b200/csrc/sampler.cu:build_padded_adj_kernel): the same table semantics as
    construct_adj (reference minibatch.py:227-245) with Philox draws instead of numpy's RandomState:
    draw j of node u = word j&3 of Philox block (counter, c2=u, BUILD tag + j>>2);
    deg < MD: MD positions mulhi32(draw_j, deg) (with replacement); deg > MD: Floyd's MD distinct positions."""
    from .sampler import _draws
    from .philox import mulhi32
    STREAM_BUILD = 0x10000000
    indptr = np.asarray(indptr).astype(np.int64)
    indices = np.asarray(indices)
    n = len(indptr) - 1
    adj = np.full((n + 1, max_degree), n, dtype=np.int32)
    deg = np.zeros((n,), dtype=np.float32)
    r_all = _draws(seed, counter, max_degree, c2=np.arange(n, dtype=np.uint32), tag=STREAM_BUILD)   # [n, MD]
    for u in range(n):
        if skip is not None and skip[u]:
            continue
        start, d = indptr[u], int(indptr[u + 1] - indptr[u])
        deg[u] = d
        if d == 0:
            continue
        if d == max_degree:
            adj[u] = indices[start:start + d]
        elif d < max_degree:
            adj[u] = indices[start + mulhi32(r_all[u], np.uint32(d)).astype(np.int64)]
        else:
            sel = []
            for j in range(max_degree):
                m = d - max_degree + j
                t = int(mulhi32(r_all[u, j], np.uint32(m + 1)))
                sel.append(m if t in sel else t)
            adj[u] = indices[start + np.array(sel, dtype=np.int64)]
    return adj, deg
