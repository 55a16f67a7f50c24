"""Backward pass of the pooling aggregators ("the three aggregators").

The gradient formulas (supervised_models.pool_branch_backward) and the autograd wiring (_PoolAggregateRowsFn: segment
offsets, accumulation over hops, scatter into the previous layer's rows) are device-agnostic torch code around the
kernels.  Here the kernels are replaced by torch stand-ins with the semantics documented in ops.py - TEST mocks only, the
product has no such path - so the whole two-layer chain can be compared with torch autograd on the oracle's op sequence
(reference aggregators.py:168-195, :246-273) without a GPU.  The GPU twin of this test drives the real kernels."""
import numpy as np
import pytest
import torch

import graphsage_b200 as gs
from graphsage_b200 import ops, supervised_models as sm


def _fake_sage_gemm(parts, combine=ops.COMBINE_ADD, bias=None, act=ops.ACT_NONE, math=None, out=None, packed=None):
    ys = [a[:, :k] @ w for (a, k, w) in parts]
    y = torch.cat(ys, dim=1) if combine == ops.COMBINE_CONCAT else sum(ys[1:], ys[0])
    if bias is not None:
        y = y + bias
    return torch.relu(y) if act == ops.ACT_RELU else y


def _fake_gather_rows(feats, ids, out=None):
    r = feats[ids.long()].float()
    if out is not None:
        out.copy_(r)
        return out
    return r


def _fake_gather_mean(src, segments, include_self=False, want_self=True, out_pitch=None, out_mean=None, out_self=None):
    (s,) = segments
    assert s.neigh_ids is None and not include_self and not want_self
    return None, src[s.neigh_row0:s.neigh_row0 + s.n * s.k].reshape(s.n, s.k, -1).mean(dim=1)


@pytest.fixture()
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(ops, "sage_gemm", _fake_sage_gemm)
    monkeypatch.setattr(ops, "gather_rows", _fake_gather_rows)
    monkeypatch.setattr(ops, "gather_mean", _fake_gather_mean)
    monkeypatch.setattr(ops, "segment_max", lambda x, n, k: x.reshape(n, k, -1).amax(dim=1))


def _ref_layer(selfv, neigh, k, w, pool, concat, last):
    """The oracle's op sequence in differentiable torch (amax splits the gradient evenly among ties, like TF)."""
    n = selfv.shape[0]
    h = torch.relu(neigh @ w["mlp_weights"] + w["mlp_bias"]).reshape(n, k, -1)
    hp = h.amax(dim=1) if pool == "max" else h.mean(dim=1)
    fs, fn = selfv @ w["self_weights"], hp @ w["neigh_weights"]
    y = torch.cat([fs, fn], dim=1) if concat else fs + fn
    return y if last else torch.relu(y)


@pytest.mark.parametrize("pool", ["max", "mean"])
@pytest.mark.parametrize("concat", [True, False])
def test_two_layer_pool_chain_gradients_match_autograd(cpu_kernels, pool, concat):
    r = np.random.RandomState(3)
    N, F, D, B, k1, k2 = 40, 10, 6, 5, 3, 4            # seeds B, hop-1 fanout k1, hop-2 fanout k2
    feats = torch.from_numpy(r.randn(N, F).astype(np.float32))
    feats[7] = 0.0                                       # an all-zero row: every MLP unit ties at relu(bias)
    s0 = torch.from_numpy(r.randint(0, N, size=B).astype(np.int32))
    s1 = torch.from_numpy(r.randint(0, N, size=B * k1).astype(np.int32))
    s2 = torch.from_numpy(r.randint(0, N, size=B * k1 * k2).astype(np.int32))
    s2[:k2] = 7                                          # one whole fanout group of identical rows -> exact ties in the max
    cls = gs.MaxPoolingAggregator if pool == "max" else gs.MeanPoolingAggregator
    dim_mult = 2 if concat else 1
    a0 = cls(F, D, act=gs.relu, concat=concat, device="cpu")
    a1 = cls(dim_mult * D, D, act=gs.identity, concat=concat, device="cpu")
    for a in (a0, a1):
        a.math = ops.MATH_FP32_SIMT
        a.mlp_layers[0].vars["bias"] = torch.from_numpy(r.randn(a.hidden_dim).astype(np.float32) * 0.1)
    params = []
    for a in (a0, a1):
        for d in (a.vars, a.mlp_layers[0].vars):
            for key in d:
                d[key] = d[key].detach().clone().requires_grad_(True)
                params.append(d[key])
    # ---- through the autograd function (layer 0: id-addressed hops; layer 1: range-addressed rows of layer 0's output)
    seg0 = [ops.Seg(B, k1, self_ids=s0, neigh_ids=s1, out_row0=0), ops.Seg(B * k1, k2, self_ids=s1, neigh_ids=s2, out_row0=B)]
    m0 = a0.mlp_layers[0].vars
    h1 = sm._PoolAggregateRowsFn.apply(a0, feats, seg0, a0.vars["self_weights"], a0.vars["neigh_weights"], m0["weights"],
                                       m0["bias"])
    assert tuple(h1.shape) == (B + B * k1, dim_mult * D)
    seg1 = [ops.Seg(B, k1, self_row0=0, neigh_row0=B, out_row0=0)]
    m1 = a1.mlp_layers[0].vars
    out = sm._PoolAggregateRowsFn.apply(a1, h1, seg1, a1.vars["self_weights"], a1.vars["neigh_weights"], m1["weights"],
                                        m1["bias"])
    R = torch.from_numpy(r.randn(*out.shape).astype(np.float32))
    (out * R).sum().backward()
    got = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    # ---- the same computation as plain differentiable torch
    w0 = dict(a0.vars, mlp_weights=m0["weights"], mlp_bias=m0["bias"])
    w1 = dict(a1.vars, mlp_weights=m1["weights"], mlp_bias=m1["bias"])
    x0, x1, x2 = feats[s0.long()], feats[s1.long()], feats[s2.long()]
    r_hop0 = _ref_layer(x0, x1, k1, w0, pool, concat, last=False)
    r_hop1 = _ref_layer(x1, x2, k2, w0, pool, concat, last=False)
    ref = _ref_layer(r_hop0, r_hop1, k1, w1, pool, concat, last=True)
    assert torch.allclose(out.detach(), ref.detach(), rtol=1e-5, atol=1e-5)
    (ref * R).sum().backward()
    for p, g in zip(params, got):
        assert p.grad is not None and torch.allclose(g, p.grad, rtol=2e-4, atol=2e-5), float((g - p.grad).abs().max())


def test_pool_branch_backward_splits_ties_evenly():
    n, k, hid, F = 2, 3, 4, 5
    xn = torch.randn(n * k, F)
    h = torch.tensor([[1., 0., 2., 0.]] * 3 + [[0., 3., 0., 0.], [5., 3., 0., 0.], [5., 1., 0., 0.]])
    hp = h.reshape(n, k, hid).amax(dim=1)
    dhp = torch.ones(n, hid)
    Wm = torch.randn(F, hid)
    dWm, dbm, dxn = sm.pool_branch_backward("max", xn, h, hp, dhp, Wm, k, True)
    # group 0: three-way ties share 1/3 each where h > 0; zero activations get nothing (ReLU); group 1: two-way ties 1/2
    assert torch.allclose(dbm, torch.tensor([3 * (1 / 3) + 2 * 0.5, 2 * 0.5, 3 * (1 / 3), 0.0]))
    assert dxn.shape == (n * k, F) and dWm.shape == (F, hid)
    _, dbm_mean, none = sm.pool_branch_backward("mean", xn, h, hp, dhp, Wm, k, False)
    assert none is None and torch.allclose(dbm_mean, (h > 0).float().sum(dim=0) / k)


def test_parameter_lists_train_the_mlp_but_decay_only_aggregator_vars():
    a = gs.MaxPoolingAggregator(8, 4, concat=True, device="cpu")
    g = gs.MeanAggregator(8, 4, concat=True, device="cpu")
    every, decayed = sm.aggregator_parameters([a, g])
    assert len(every) == 2 + 2 + 2 and len(decayed) == 4
    ids = {id(t) for t in decayed}
    assert id(a.mlp_layers[0].vars["weights"]) not in ids and id(a.vars["neigh_weights"]) in ids
