"""CPU: the oracle restatement vs. the vectors produced by the reference's own
python under the TF shim (tests/golden/make_golden.py)."""
import numpy as np

import oracle
from conftest import load_golden, rel_err


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10
    def kat(c, k):
        return [int(x) for x in oracle.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))]
    assert kat([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sampler_matches_reference_under_shim():
    g = load_golden("sampler")
    for ci in range(int(g["n_cases"])):
        out = oracle.sample_padded(g["adj%d" % ci], g["ids%d" % ci], int(g["k%d" % ci]),
                                   int(g["seed%d" % ci]), int(g["counter%d" % ci]))
        assert out.dtype == np.int32
        np.testing.assert_array_equal(out, g["out%d" % ci])
        # invariants from SURVEY section 4
        n = g["adj%d" % ci].shape[0] - 1
        assert (out[g["ids%d" % ci] == n] == n).all()


def test_perm_prefix_is_prefix_of_full_perm_and_distinct():
    full = oracle.perm_prefix(9, 4, 128, 128)
    assert sorted(full) == list(range(128))
    for k in (0, 1, 10, 25, 127):
        np.testing.assert_array_equal(oracle.perm_prefix(9, 4, 128, k), full[:k])


def test_aggregators_match_reference_under_shim():
    g = load_golden("aggregators")
    s, n = g["self"], g["neigh"]
    for c in (0, 1):
        y = oracle.mean_aggregator(s, n, g["mean_c%d_nw" % c], g["mean_c%d_sw" % c], concat=bool(c))
        assert rel_err(y, g["mean_c%d_out" % c]) < 1e-6
        y = oracle.maxpool_aggregator(s, n, g["maxpool_c%d_mw" % c], g["maxpool_c%d_mb" % c],
                                      g["maxpool_c%d_nw" % c], g["maxpool_c%d_sw" % c], concat=bool(c))
        assert int(g["maxpool_c%d_hidden" % c]) == 512
        assert rel_err(y, g["maxpool_c%d_out" % c]) < 1e-6
    assert rel_err(oracle.gcn_aggregator(s, n, g["gcn_w"]), g["gcn_out"]) < 1e-6
    y = oracle.mean_aggregator(s, g["neigh2"], g["mean_id_nw"], g["mean_id_sw"], concat=True, act=lambda x: x)
    assert rel_err(y, g["mean_id_out"]) < 1e-6
    assert (y < 0).any()                       # identity activation really is identity
    r = oracle.glorot_range((50, 16))
    assert np.abs(g["glorot_sample"]).max() <= r and np.abs(g["glorot_sample"]).max() > 0.8 * r


def _aggs(g, model, L):
    kind = {"mean": "mean", "mean3": "mean", "gcn": "gcn", "maxpool": "maxpool"}[model]
    out = []
    for li in range(L):
        d = {"type": kind}
        for key in ("neigh_weights", "self_weights", "weights", "mlp_weights", "mlp_bias"):
            name = "%s_L%d_%s" % (model, li, key)
            if name in g:
                d[key] = g[name]
        out.append(d)
    return out


def test_khop_matches_reference_under_shim():
    g = load_golden("khop")
    for model in ("mean", "gcn", "maxpool", "mean3"):
        fan = [int(x) for x in g[model + "_fanout"]]
        L = len(fan)
        samples, support = oracle.sample_khop(g["adj"], g["seeds"], fan, 123, 40)
        assert support == [int(x) for x in g[model + "_support"]]
        for h in range(L + 1):
            np.testing.assert_array_equal(samples[h], g["%s_samples%d" % (model, h)])
        out = oracle.aggregate_khop(samples, g["feats"], fan, support, len(g["seeds"]), _aggs(g, model, L),
                                    bool(g[model + "_concat"]))
        assert rel_err(out, g[model + "_out"]) < 1e-5
        assert rel_err(oracle.l2_normalize(out), g[model + "_out_l2"]) < 1e-5
        out2 = oracle.forward_2hop(g["adj"], g["feats"], g["seeds"], fan, _aggs(g, model, L),
                                   bool(g[model + "_concat"]), 123, 40, normalize=True)
        assert rel_err(out2, g[model + "_out_l2"]) < 1e-5


def test_padded_adjacency_matches_reference_under_shim():
    g = load_golden("adjacency")
    order = [int(x) for x in g["node_order"]]
    ptr = g["nb_ptr"]
    nbrs, removed, vt = {}, {}, {}
    for pos, u in enumerate(order):
        nb = [int(v) for v in g["nb_idx"][ptr[pos]:ptr[pos + 1]]]
        nbrs[u] = nb
        vt[u] = bool(g["val_or_test"][pos])
        for j, v in enumerate(nb):
            removed[(u, v)] = bool(g["nb_removed"][ptr[pos] + j])
    id2idx = {u: u for u in order}
    rng = np.random.RandomState(123)
    adj, deg = oracle.construct_adj(order, nbrs, id2idx, vt, removed, int(g["max_degree"]), rng)
    test_adj = oracle.construct_test_adj(order, nbrs, id2idx, int(g["max_degree"]), rng)
    np.testing.assert_array_equal(adj, g["adj"])
    np.testing.assert_array_equal(deg, g["deg"])
    np.testing.assert_array_equal(test_adj, g["test_adj"])
    n = len(order)
    assert (adj[n] == n).all() and adj.dtype == np.int32


def test_sample_csr_properties():
    rs = np.random.RandomState(0)
    n = 40
    deg = rs.randint(0, 30, size=n)
    deg[:3] = [0, 1, 10]
    indptr = np.concatenate([[0], np.cumsum(deg)])
    indices = np.concatenate([rs.choice(1000, d, replace=False) for d in deg]).astype(np.int32)
    ids = np.arange(n, dtype=np.int32)
    k = 10
    out = oracle.sample_csr(indptr, indices, ids, k, 7, 3, True, pad_id=-1)
    for i in range(n):
        nb = set(indices[indptr[i]:indptr[i + 1]].tolist())
        if deg[i] == 0:
            assert (out[i] == -1).all()
        else:
            assert set(out[i].tolist()) <= nb
            if deg[i] >= k:
                assert len(set(out[i].tolist())) == k       # without replacement
    out2 = oracle.sample_csr(indptr, indices, ids, k, 7, 3, False, pad_id=-1)
    i = 1
    assert out2[i, 0] == indices[indptr[i]] and (out2[i, 1:] == -1).all()
    # different counters give different draws, same counter reproduces
    np.testing.assert_array_equal(out, oracle.sample_csr(indptr, indices, ids, k, 7, 3, True, pad_id=-1))
    assert (out != oracle.sample_csr(indptr, indices, ids, k, 7, 4, True, pad_id=-1)).any()


def test_torch_cpu_baseline_matches_numpy_oracle():
    import torch
    from oracle import torch_ref
    g = load_golden("khop")
    for model, kind in (("mean", "mean"), ("gcn", "gcn"), ("maxpool", "maxpool")):
        fan = [int(x) for x in g[model + "_fanout"]]
        aggs = [{k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in a.items() if k != "type"}
                for a in _aggs(g, model, len(fan))]
        out = torch_ref.forward(torch.from_numpy(g["adj"]), torch.from_numpy(g["feats"]), torch.from_numpy(g["seeds"]),
                                fan, aggs, bool(g[model + "_concat"]), kind, 123, 40, normalize=True)
        assert rel_err(out.numpy(), g[model + "_out_l2"]) < 1e-5


def test_unigram_sampler_oracle_distribution():
    rs = np.random.RandomState(0)
    deg = rs.randint(0, 100, size=1000).astype(np.float64)
    x = oracle.sample_unigram(deg, 20, 123, 5)
    assert x.dtype == np.int32 and (deg[x] > 0).all()
    np.testing.assert_array_equal(x, oracle.sample_unigram(deg, 20, 123, 5))
    assert (x != oracle.sample_unigram(deg, 20, 123, 6)).any()
    big = oracle.sample_unigram(deg, 200000, 1, 1)
    p = np.bincount(big, minlength=1000) / 200000.0
    q = deg ** 0.75 / (deg ** 0.75).sum()
    assert np.abs(p - q).max() < 1e-3


def test_meanpool_matches_reference_under_shim():
    g = load_golden("meanpool")
    for c in (0, 1):
        y = oracle.meanpool_aggregator(g["self"], g["neigh"], g["c%d_mw" % c], g["c%d_mb" % c], g["c%d_nw" % c],
                                       g["c%d_sw" % c], concat=bool(c))
        assert rel_err(y, g["c%d_out" % c]) < 1e-6


def test_build_padded_adj_oracle_properties():
    rs = np.random.RandomState(3)
    n, md = 200, 8
    deg = rs.randint(0, 30, size=n)
    deg[:3] = [0, md, md + 5]
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = np.concatenate([rs.choice(n, d, replace=False) for d in deg]).astype(np.int32)
    adj, d = oracle.build_padded_adj(indptr, indices, md, 5, 0)
    assert adj.shape == (n + 1, md) and (adj[n] == n).all() and (adj[0] == n).all()
    np.testing.assert_array_equal(adj[1], indices[indptr[1]:indptr[2]])
    assert len(set(adj[2].tolist())) == md and set(adj[2].tolist()) <= set(indices[indptr[2]:indptr[3]].tolist())
    np.testing.assert_array_equal(d, deg.astype(np.float32))
