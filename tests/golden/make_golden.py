"""Generate tests/golden/*.npz by executing the reference's own hot-path python
(/root/reference/graphsage/{neigh_samplers,aggregators,layers,inits,models,minibatch}.py)
under the numpy TF shim (tf_shim.py).  Run HERE (the container that has
/root/reference); the GPU box only reads the committed .npz files.

    python tests/golden/make_golden.py

Nothing from /root/reference is copied: the modules are imported from where they lie.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import tf_shim  # noqa: E402

tf = tf_shim.install()

from graphsage.neigh_samplers import UniformNeighborSampler  # noqa: E402
from graphsage.aggregators import MeanAggregator, GCNAggregator, MaxPoolingAggregator, MeanPoolingAggregator  # noqa: E402
from graphsage.models import SampleAndAggregate, SAGEInfo  # noqa: E402
from graphsage.minibatch import NodeMinibatchIterator  # noqa: E402
from graphsage.inits import glorot  # noqa: E402

rs = np.random.RandomState(7)


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, {k: np.asarray(v).shape for k, v in kw.items()})


# ------------------------------------------------------------------ sampler
def make_adj(n, md):
    adj = rs.randint(0, n, size=(n + 1, md)).astype(np.int32)
    adj[n, :] = n                       # dummy row (minibatch.py:228)
    adj[3, :] = n                       # an isolated node
    return adj


def golden_sampler():
    cases = {}
    for ci, (n, md, nid, k, seed, counter) in enumerate(
            [(50, 16, 23, 5, 123, 0), (200, 128, 64, 25, 123, 1), (200, 128, 7, 10, 99, 1 << 33), (30, 8, 11, 8, 5, 2)]):
        adj = make_adj(n, md)
        ids = rs.randint(0, n + 1, size=nid).astype(np.int32)
        ids[0] = n                      # dummy id -> all outputs n
        tf_shim.SHUFFLE_SEED, tf_shim.SHUFFLE_COUNTER = seed, counter
        sampler = UniformNeighborSampler(adj)
        out = sampler((ids, k))
        cases.update({"adj%d" % ci: adj, "ids%d" % ci: ids, "k%d" % ci: k, "seed%d" % ci: seed,
                      "counter%d" % ci: np.uint64(counter), "out%d" % ci: out.astype(np.int32)})
    cases["n_cases"] = ci + 1
    save("sampler", **cases)


# ------------------------------------------------------------------ aggregators
def golden_aggregators():
    n, k, din, dout = 37, 10, 50, 16
    selfv = rs.randn(n, din).astype(np.float32)
    neigh = rs.randn(n, k, din).astype(np.float32)
    neigh[5] = 0.0                      # a node whose neighbours are all the dummy zero row
    out = {"self": selfv, "neigh": neigh}
    for concat in (False, True):
        agg = MeanAggregator(din, dout, concat=concat)
        tag = "mean_c%d" % concat
        out[tag + "_nw"], out[tag + "_sw"] = agg.vars["neigh_weights"], agg.vars["self_weights"]
        out[tag + "_out"] = agg((selfv, neigh))
        agg = MaxPoolingAggregator(din, dout, concat=concat)
        tag = "maxpool_c%d" % concat
        out[tag + "_nw"], out[tag + "_sw"] = agg.vars["neigh_weights"], agg.vars["self_weights"]
        out[tag + "_mw"], out[tag + "_mb"] = agg.mlp_layers[0].vars["weights"], agg.mlp_layers[0].vars["bias"] + \
            rs.randn(agg.hidden_dim).astype(np.float32) * 0.1
        agg.mlp_layers[0].vars["bias"] = out[tag + "_mb"]     # non-zero bias so the bias add is exercised
        out[tag + "_hidden"] = agg.hidden_dim
        out[tag + "_out"] = agg((selfv, neigh))
    agg = GCNAggregator(din, dout)
    out["gcn_w"] = agg.vars["weights"]
    out["gcn_out"] = agg((selfv, neigh))
    # identity activation (last layer, models.py:307-310) and distinct neigh_input_dim
    neigh2 = rs.randn(n, k, 24).astype(np.float32)
    agg = MeanAggregator(din, dout, neigh_input_dim=24, act=lambda x: x, concat=True)
    out["mean_id_nw"], out["mean_id_sw"], out["neigh2"] = agg.vars["neigh_weights"], agg.vars["self_weights"], neigh2
    out["mean_id_out"] = agg((selfv, neigh2))
    out["glorot_sample"] = glorot([50, 16])
    save("aggregators", **out)


# ------------------------------------------------------------------ K-hop recursion
class _Stub(object):
    """Just the attributes SampleAndAggregate.sample/.aggregate read (models.py:254-330)."""


def golden_khop():
    out = {}
    n, md, f = 300, 32, 20
    adj = make_adj(n, md)
    feats = np.vstack([rs.randn(n, f).astype(np.float32), np.zeros((1, f), np.float32)])   # supervised_train.py:133-135
    B = 9
    seeds = rs.randint(0, n, size=B).astype(np.int32)
    out.update(adj=adj, feats=feats, seeds=seeds)
    for model, cls, concat, dims, fan in [("mean", MeanAggregator, True, [f, 12, 8], [5, 3]),
                                          ("gcn", GCNAggregator, False, [f, 16, 16], [5, 3]),
                                          ("maxpool", MaxPoolingAggregator, True, [f, 12, 8], [4, 2]),
                                          ("mean3", MeanAggregator, True, [f, 8, 8, 6], [4, 3, 2])]:
        tf_shim.SHUFFLE_SEED, tf_shim.SHUFFLE_COUNTER = 123, 40
        sampler = UniformNeighborSampler(adj)
        infos = [SAGEInfo("node", sampler, fan[i], dims[i + 1]) for i in range(len(fan))]
        stub = _Stub()
        stub.batch_size = B
        stub.aggregator_cls = cls
        stub.placeholders = {"dropout": 0.0}
        samples, support = SampleAndAggregate.sample(stub, seeds, infos)
        hidden, aggs = SampleAndAggregate.aggregate(stub, samples, [feats][0], dims, fan, support, concat=concat)
        for h, s in enumerate(samples):
            out["%s_samples%d" % (model, h)] = np.asarray(s).astype(np.int32)
        out[model + "_support"] = np.array(support)
        out[model + "_fanout"] = np.array(fan)
        out[model + "_dims"] = np.array(dims)
        out[model + "_concat"] = concat
        out[model + "_out"] = hidden
        out[model + "_out_l2"] = tf.nn.l2_normalize(hidden, 1)          # models.py:368
        for li, a in enumerate(aggs):
            for key, v in a.vars.items():
                out["%s_L%d_%s" % (model, li, key)] = v
            if hasattr(a, "mlp_layers"):
                out["%s_L%d_mlp_weights" % (model, li)] = a.mlp_layers[0].vars["weights"]
                out["%s_L%d_mlp_bias" % (model, li)] = a.mlp_layers[0].vars["bias"]
    save("khop", **out)


# ------------------------------------------------------------------ padded adjacency
class _FakeG(object):
    """networkx-1.11-shaped view (G.nodes(), G.node[n], G.neighbors(n), G[u][v]) over plain dicts."""

    def __init__(self, nodes, nbrs, attrs, eattrs):
        self._nodes, self._nbrs, self.node, self._e = nodes, nbrs, attrs, eattrs

    def nodes(self):
        return list(self._nodes)

    def neighbors(self, n):
        return list(self._nbrs[n])

    def __getitem__(self, u):
        return {v: self._e[(u, v)] for v in self._nbrs[u]}


def golden_adjacency():
    n, md = 60, 8
    nodes = ["n%d" % i for i in rs.permutation(n)]
    id2idx = {"n%d" % i: i for i in range(n)}
    nbrs = {u: [] for u in nodes}
    for i in range(n):
        deg = [0, 1, 3, 8, 12, 20][rs.randint(6)]
        for j in rs.choice(n, size=deg, replace=False):
            u, v = "n%d" % i, "n%d" % j
            if u != v and v not in nbrs[u]:
                nbrs[u].append(v)
                nbrs[v].append(u)
    attrs = {u: {"val": bool(rs.rand() < 0.1), "test": bool(rs.rand() < 0.15)} for u in nodes}
    eattrs = {}
    for u in nodes:
        for v in nbrs[u]:
            eattrs[(u, v)] = {"train_removed": attrs[u]["val"] or attrs[u]["test"] or attrs[v]["val"] or attrs[v]["test"]}
    G = _FakeG(nodes, nbrs, attrs, eattrs)
    np.random.seed(123)
    it = NodeMinibatchIterator(G, id2idx, None, {u: 0 for u in nodes}, 2, batch_size=4, max_degree=md)
    flat_nb = np.array([id2idx[v] for u in nodes for v in nbrs[u]], dtype=np.int32)
    nb_ptr = np.cumsum([0] + [len(nbrs[u]) for u in nodes]).astype(np.int64)
    flat_removed = np.array([eattrs[(u, v)]["train_removed"] for u in nodes for v in nbrs[u]], dtype=bool)
    save("adjacency", node_order=np.array([id2idx[u] for u in nodes], dtype=np.int32), nb_ptr=nb_ptr, nb_idx=flat_nb,
         nb_removed=flat_removed, val_or_test=np.array([attrs[u]["val"] or attrs[u]["test"] for u in nodes]),
         max_degree=md, adj=it.adj.astype(np.int32), deg=it.deg, test_adj=it.test_adj.astype(np.int32))


def golden_meanpool():
    """MeanPoolingAggregator (reference aggregators.py:197-273), generated separately so the older fixtures keep their bytes."""
    rs2 = np.random.RandomState(21)
    n, k, din, dout = 29, 7, 40, 16
    selfv = rs2.randn(n, din).astype(np.float32)
    neigh = rs2.randn(n, k, din).astype(np.float32)
    out = {"self": selfv, "neigh": neigh}
    for concat in (False, True):
        agg = MeanPoolingAggregator(din, dout, concat=concat)
        tag = "c%d" % concat
        agg.mlp_layers[0].vars["bias"] = agg.mlp_layers[0].vars["bias"] + rs2.randn(agg.hidden_dim).astype(np.float32) * 0.1
        out[tag + "_nw"], out[tag + "_sw"] = agg.vars["neigh_weights"], agg.vars["self_weights"]
        out[tag + "_mw"], out[tag + "_mb"] = agg.mlp_layers[0].vars["weights"], agg.mlp_layers[0].vars["bias"]
        out[tag + "_out"] = agg((selfv, neigh))
    save("meanpool", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "meanpool":
        golden_meanpool()
        sys.exit(0)
    golden_sampler()
    golden_aggregators()
    golden_khop()
    golden_adjacency()
    golden_meanpool()
