"""GPU: the training step - loss and gradients of SupervisedGraphsage against torch-CPU autograd
on the oracle's op sequence (oracle/torch_ref.py restates reference models.py:254-330 + supervised_models.py:78-126)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _cpu_loss(adj, feats, seeds, labels, fan, aggs, head, concat, kind, seed, counter, sigmoid, wd):
    out = torch_ref.forward(torch.from_numpy(adj), torch.from_numpy(feats), torch.from_numpy(seeds), fan, aggs, concat,
                            kind, seed, counter, normalize=True)
    logits = out @ head["weights"] + head["bias"]
    loss = torch.zeros(())
    if wd:
        for a in aggs:
            for v in a.values():
                loss = loss + wd * 0.5 * (v * v).sum()
        for v in head.values():
            loss = loss + wd * 0.5 * (v * v).sum()
    y = torch.from_numpy(labels)
    if sigmoid:
        loss = loss + torch.nn.functional.binary_cross_entropy_with_logits(logits, y, reduction="mean")
    else:
        loss = loss + (-(y * torch.log_softmax(logits, dim=1)).sum(dim=1)).mean()
    return loss


@pytest.mark.parametrize("kind,concat,sigmoid", [("mean", True, True), ("mean", False, False), ("gcn", False, True)])
def test_loss_and_gradients_match_cpu_autograd(kind, concat, sigmoid):
    import graphsage_b200 as gs
    g = load_golden("khop")
    rs = np.random.RandomState(5)
    adj, feats = g["adj"], g["feats"]
    n, B, C = adj.shape[0] - 1, 24, 7
    seeds = rs.randint(0, n, size=B).astype(np.int32)
    labels = (rs.rand(B, C) < 0.3).astype(np.float32) if sigmoid else np.eye(C, dtype=np.float32)[rs.randint(0, C, size=B)]
    fan, dim, wd = [5, 3], 12, 1e-3
    gs.set_default_math("fp32")
    sampler = gs.UniformNeighborSampler(torch.from_numpy(adj).cuda(), seed=123)
    sampler.counter = 40
    infos = [gs.SAGEInfo("node", sampler, fan[0], dim), gs.SAGEInfo("node", sampler, fan[1], dim)]
    m = gs.SupervisedGraphsage(C, {"batch_size": B, "dropout": 0.}, torch.from_numpy(feats).cuda(),
                               torch.from_numpy(adj).cuda(), None, infos, concat=concat, aggregator_type=kind,
                               sigmoid_loss=sigmoid, learning_rate=0.01, weight_decay=wd)
    # the same parameters on the CPU side, as autograd leaves
    aggs = []
    for a in m.aggregators:
        aggs.append({k: v.detach().cpu().clone().requires_grad_(True) for k, v in a.vars.items()})
    head = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.node_pred_vars.items()}
    ref = _cpu_loss(adj, feats, seeds, labels, fan, aggs, head, concat, kind, 123, 40, sigmoid, wd)
    ref.backward()
    loss = m.loss(torch.from_numpy(seeds), torch.from_numpy(labels))
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    for a, ra in zip(m.aggregators, aggs):
        for k in a.vars:
            assert rel_err(a.vars[k].grad.cpu().numpy(), ra[k].grad.numpy(), floor=1e-8) < 2e-4, (kind, k)
    for k in head:
        assert rel_err(m.node_pred_vars[k].grad.cpu().numpy().reshape(1, -1), head[k].grad.numpy().reshape(1, -1)) < 2e-4


def test_training_steps_track_cpu_adam():
    """Five clipped-Adam steps on the GPU path follow the same steps on the CPU restatement."""
    import graphsage_b200 as gs
    g = load_golden("khop")
    rs = np.random.RandomState(9)
    adj, feats = g["adj"], g["feats"]
    n, B, C = adj.shape[0] - 1, 32, 5
    fan, dim = [5, 3], 16
    gs.set_default_math("fp32")
    sampler = gs.UniformNeighborSampler(torch.from_numpy(adj).cuda(), seed=7)
    infos = [gs.SAGEInfo("node", sampler, fan[0], dim), gs.SAGEInfo("node", sampler, fan[1], dim)]
    m = gs.SupervisedGraphsage(C, {"batch_size": B, "dropout": 0.}, torch.from_numpy(feats).cuda(),
                               torch.from_numpy(adj).cuda(), None, infos, concat=True, aggregator_type="mean",
                               sigmoid_loss=False, learning_rate=0.01, weight_decay=0.0)
    aggs = [{k: v.detach().cpu().clone().requires_grad_(True) for k, v in a.vars.items()} for a in m.aggregators]
    head = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.node_pred_vars.items()}
    params = [v for a in aggs for v in a.values()] + list(head.values())
    opt = torch.optim.Adam(params, lr=0.01)
    w_true = rs.randn(feats.shape[1], C).astype(np.float32)
    gpu_losses, cpu_losses = [], []
    for step in range(5):
        seeds = rs.randint(0, n, size=B).astype(np.int32)
        labels = np.eye(C, dtype=np.float32)[(feats[seeds] @ w_true).argmax(1)]
        gpu_losses.append(float(m.train_step(torch.from_numpy(seeds), torch.from_numpy(labels))))
        opt.zero_grad()
        ref = _cpu_loss(adj, feats, seeds, labels, fan, aggs, head, True, "mean", 7, 2 * step, False, 0.0)
        ref.backward()
        for p in params:
            p.grad.clamp_(-5.0, 5.0)
        opt.step()
        cpu_losses.append(float(ref))
    assert np.allclose(gpu_losses, cpu_losses, rtol=2e-3), (gpu_losses, cpu_losses)
    for a, ra in zip(m.aggregators, aggs):
        for k in a.vars:
            assert rel_err(a.vars[k].detach().cpu().numpy(), ra[k].detach().numpy()) < 5e-3


def test_unigram_sampler_bit_exact_and_distribution():
    import graphsage_b200 as gs
    import oracle
    rs = np.random.RandomState(2)
    deg = rs.randint(0, 300, size=5000).astype(np.float64)
    deg[:5] = 0
    s = gs.UnigramNegativeSampler(deg, seed=9)
    a = s(20).cpu().numpy()
    b = s(1000).cpu().numpy()
    np.testing.assert_array_equal(a, oracle.sample_unigram(deg, 20, 9, 0))
    np.testing.assert_array_equal(b, oracle.sample_unigram(deg, 1000, 9, 1))
    assert (deg[b] > 0).all()                                    # zero-degree nodes are never drawn
    big = gs.ops.sample_unigram(s.cdf, 400000, 1, 1).cpu().numpy()
    p = np.bincount(big, minlength=5000) / 400000.0
    q = deg ** 0.75 / (deg ** 0.75).sum()
    assert np.abs(p - q).max() < 5e-4


def test_unsupervised_loss_gradients_and_mrr_match_cpu():
    import graphsage_b200 as gs
    import oracle
    g = load_golden("khop")
    rs = np.random.RandomState(11)
    adj, feats = g["adj"], g["feats"]
    n, B, NEG = adj.shape[0] - 1, 16, 20
    deg = rs.randint(1, 40, size=n).astype(np.float64)
    b1 = rs.randint(0, n, size=B).astype(np.int32)
    b2 = rs.randint(0, n, size=B).astype(np.int32)
    fan, dim = [5, 3], 12
    gs.set_default_math("fp32")
    sampler = gs.UniformNeighborSampler(torch.from_numpy(adj).cuda(), seed=123)
    infos = [gs.SAGEInfo("node", sampler, fan[0], dim), gs.SAGEInfo("node", sampler, fan[1], dim)]
    m = gs.UnsupervisedGraphsage({"batch_size": B, "dropout": 0.}, torch.from_numpy(feats).cuda(),
                                 torch.from_numpy(adj).cuda(), deg, infos, concat=True, aggregator_type="mean",
                                 neg_sample_size=NEG, learning_rate=0.01, weight_decay=1e-3, seed=77)
    aggs = [{k: v.detach().cpu().clone().requires_grad_(True) for k, v in a.vars.items()} for a in m.aggregators]
    # CPU restatement: the three passes use sampler counters 0..1, 2..3, 4..5 (two calls each), negatives from counter 0
    neg = oracle.sample_unigram(deg, NEG, 77, 0)
    A, F = torch.from_numpy(adj), torch.from_numpy(feats)
    o1 = torch_ref.forward(A, F, torch.from_numpy(b1), fan, aggs, True, "mean", 123, 0)
    o2 = torch_ref.forward(A, F, torch.from_numpy(b2), fan, aggs, True, "mean", 123, 2)
    on = torch_ref.forward(A, F, torch.from_numpy(neg), fan, aggs, True, "mean", 123, 4)
    aff = (o1 * o2).sum(1)
    neg_aff = o1 @ on.t()
    ref = torch.nn.functional.softplus(-aff).sum() + torch.nn.functional.softplus(neg_aff).sum()
    for a in aggs:
        for v in a.values():
            ref = ref + 1e-3 * 0.5 * (v * v).sum()
    ref = ref / B
    ref.backward()
    loss = m.loss(torch.from_numpy(b1), torch.from_numpy(b2))
    loss.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < 1e-5 * max(1.0, abs(float(ref.detach())))
    for a, ra in zip(m.aggregators, aggs):
        for k in a.vars:
            assert rel_err(a.vars[k].grad.cpu().numpy(), ra[k].grad.numpy(), floor=1e-8) < 3e-4, k
    aff_all = torch.cat([neg_aff.detach(), aff.detach().unsqueeze(1)], dim=1)
    ranks = torch.argsort(torch.argsort(aff_all, dim=1, descending=True, stable=True), dim=1, stable=True)
    assert abs(float(m.mrr()) - float((1.0 / (ranks[:, -1] + 1).float()).mean())) < 1e-6
    l0 = float(m.train_step(torch.from_numpy(b1), torch.from_numpy(b2)))
    assert np.isfinite(l0)
