"""A small undirected graph with the networkx-1.11 call surface the reference's data path touches
(reference graphsage/utils.py:19-75, graphsage/minibatch.py): `G.nodes()`, `G.node[n]`, `G.neighbors(n)`,
`G[u][v]`, `G.edges()`, `G.degree(n)`, `G.remove_node(n)`, `G.subgraph(nodes)`, and `node_link_graph` for the
`<prefix>-G.json` files.  The reference pins networkx <= 1.11 (utils.py:11-14), which is not installable here and whose
`G.node` / list-returning `G.nodes()` API is gone from current releases - hence this stand-in.

Iteration orders are insertion orders (nodes in JSON order, a node's neighbours in link order), which is what a
networkx-1.11 graph gives on a Python whose dicts keep insertion order.

Host-side, start-up-time code: the hot path only ever sees the int32 tables built from it (`to_csr`, minibatch.py).
"""
import numpy as np


class Graph(object):
    def __init__(self):
        self.node = {}        # node -> attribute dict
        self._adj = {}        # node -> {neighbour: edge attribute dict (one dict shared by both directions)}
        self.graph = {}

    # ---- construction
    def add_node(self, n, **attr):
        if n not in self.node:
            self.node[n] = {}
            self._adj[n] = {}
        self.node[n].update(attr)

    def add_edge(self, u, v, **attr):
        self.add_node(u)
        self.add_node(v)
        data = self._adj[u].get(v, {})
        data.update(attr)
        self._adj[u][v] = data
        self._adj[v][u] = data

    def remove_node(self, n):
        for v in list(self._adj[n]):
            if v != n:
                del self._adj[v][n]
        del self._adj[n]
        del self.node[n]

    # ---- queries (lists, as networkx 1.x returns)
    def nodes(self):
        return list(self.node)

    def neighbors(self, n):
        return list(self._adj[n])

    def degree(self, n):
        return len(self._adj[n]) + (1 if n in self._adj[n] else 0)    # a self loop counts twice

    def edges(self):
        """Each undirected edge once, (u, v) with u the endpoint met first in node order."""
        seen = set()
        out = []
        for u, nbrs in self._adj.items():
            for v in nbrs:
                if v not in seen:
                    out.append((u, v))
            seen.add(u)
        return out

    def number_of_nodes(self):
        return len(self.node)

    def number_of_edges(self):
        return len(self.edges())

    def subgraph(self, nodes):
        keep = set(nodes)
        H = Graph()
        for n in self.node:
            if n in keep:
                H.add_node(n, **self.node[n])
        for u, v in self.edges():
            if u in keep and v in keep:
                H.add_edge(u, v, **self._adj[u][v])
        return H

    def __getitem__(self, n):
        return self._adj[n]

    def __contains__(self, n):
        return n in self.node

    def __len__(self):
        return len(self.node)

    def __iter__(self):
        return iter(self.node)


def node_link_graph(data):
    """networkx 1.x `json_graph.node_link_graph` for undirected simple graphs: node attribute `id` names the node,
    a link's `source` / `target` are POSITIONS in the node list (the 1.x convention the reference's files use)."""
    if data.get("directed", False) or data.get("multigraph", False):
        raise ValueError("node_link_graph: only undirected simple graphs (the reference's datasets) are supported")
    G = Graph()
    G.graph = dict(data.get("graph", {})) if isinstance(data.get("graph", {}), dict) else dict(data.get("graph", []))
    mapping = []
    for d in data["nodes"]:
        attr = dict(d)
        n = attr.pop("id")
        mapping.append(n)
        G.add_node(n, **attr)
    for d in data["links"]:
        attr = dict(d)
        u, v = mapping[attr.pop("source")], mapping[attr.pop("target")]
        G.add_edge(u, v, **attr)
    return G


def to_csr(G, id2idx, edge_flag="train_removed"):
    """Index-space view of G for the table builders (minibatch.construct_adj, gs_build_padded_adj):
    returns dict(indptr int64 [N+1], indices int32 [nnz], edge_removed bool [nnz], val_or_test bool [N],
    node_order int32 [len(G)]) where N = len(id2idx), row id2idx[u] lists id2idx[v] for v in G.neighbors(u) in order."""
    n = len(id2idx)
    order = np.array([id2idx[u] for u in G.nodes()], dtype=np.int32)
    counts = np.zeros(n, dtype=np.int64)
    for u in G.nodes():
        counts[id2idx[u]] = len(G[u])
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = np.zeros(int(indptr[-1]), dtype=np.int32)
    removed = np.zeros(int(indptr[-1]), dtype=bool)
    vt = np.zeros(n, dtype=bool)
    for u in G.nodes():
        iu = id2idx[u]
        a = G.node[u]
        vt[iu] = bool(a.get("val", False) or a.get("test", False))
        p = indptr[iu]
        for v, e in G[u].items():
            indices[p] = id2idx[v]
            removed[p] = bool(e.get(edge_flag, False))
            p += 1
    return dict(indptr=indptr, indices=indices, edge_removed=removed, val_or_test=vt, node_order=order)
