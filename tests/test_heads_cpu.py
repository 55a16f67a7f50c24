"""Loss heads around the hot path's output against fixtures written by the reference's own
prediction.py / models.py:_accuracy / supervised_models.py:_loss under the numpy TF shim (tests/golden/heads.npz,
tests/golden/make_golden.py:golden_heads).  Torch on CPU: the heads are device-agnostic torch code."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from graphsage_b200 import supervised_models, unsupervised_models
from graphsage_b200.prediction import BipartiteEdgePredLayer, mrr_from_affinities

TOL = 2e-6


def close(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.fixture(scope="module")
def g():
    return load_golden("heads")


def t(x):
    return torch.from_numpy(np.asarray(x))


def test_edge_pred_layer_losses_match_reference(g):
    o1, o2, on = t(g["o1"]), t(g["o2"]), t(g["on"])
    for fn in ("xent", "skipgram", "hinge"):
        layer = BipartiteEdgePredLayer(16, 16, {"dropout": 0.}, loss_fn=fn, device="cpu", name="edge_predict")
        assert close(float(layer.loss(o1, o2, on)), float(g["loss_" + fn])), fn
    assert close(layer.affinity(o1, o2).numpy(), g["aff"]) and close(layer.neg_cost(o1, on).numpy(), g["neg_aff"])
    assert layer.vars == {} and layer.output_dim == 1 and layer.margin == 0.1
    with pytest.raises(ValueError):
        BipartiteEdgePredLayer(16, 16, None, loss_fn="nope", device="cpu")
    with pytest.raises(AssertionError):
        BipartiteEdgePredLayer(16, 16, None, device="cpu", bogus=1)             # Layer's kwarg whitelist (layers.py:43-45)


def test_edge_pred_layer_bilinear_matches_reference(g):
    o1, o2, on = t(g["o1"]), t(g["o2"]), t(g["on"])
    layer = BipartiteEdgePredLayer(16, 16, None, neg_sample_weights=0.25, bilinear_weights=True, device="cpu")
    assert tuple(layer.vars["weights"].shape) == (16, 16)
    r = float(np.sqrt(6.0 / 32))
    assert float(layer.vars["weights"].abs().max()) <= r                        # xavier-uniform range
    layer.vars["weights"] = t(g["bil_w"])
    assert close(layer.affinity(o1, o2).numpy(), g["bil_aff"]) and close(layer.neg_cost(o1, on).numpy(), g["bil_neg_aff"])
    assert close(float(layer.loss(o1, o2, on)), float(g["bil_loss"]))


def test_mrr_matches_reference(g):
    aff, neg_aff = t(g["aff"]), t(g["neg_aff"])
    assert close(float(mrr_from_affinities(aff, neg_aff)), float(g["mrr"]))
    # the reference's `ranks` tensor: rank of every column; the true pair is the last one
    aff_all = torch.cat([neg_aff, aff.unsqueeze(1)], dim=1)
    order = torch.argsort(aff_all, dim=1, descending=True, stable=True)
    ranks = torch.argsort(order, dim=1, stable=True)
    assert np.array_equal(ranks.numpy(), g["ranks"])
    assert int(ranks[3, -1]) == 0                                               # the planted identical pair ranks first


def test_supervised_loss_and_predictions_match_reference(g):
    logits = t(g["logits"])
    params = [t(g["a0_nw"]), t(g["a0_sw"]), t(g["a1_w"]), t(g["head_w"]), t(g["head_b"])]
    for sig, labels, tag in ((True, g["multi"], "sig"), (False, g["onehot"], "soft")):
        base = supervised_models.classification_loss(logits, t(labels), sig)
        assert close(float(base), float(g["sup_%s_wd0" % tag]))
        full = base + supervised_models.weight_decay_term(params, 0.05)
        assert close(float(full), float(g["sup_%s_wd1" % tag]))
    assert close(torch.sigmoid(logits).numpy(), g["pred_sig"]) and close(torch.softmax(logits, dim=1).numpy(), g["pred_soft"])


def test_model_methods_route_through_the_heads(g):
    """The model classes' loss()/mrr()/predict() with the hot path stubbed out (it needs a GPU): the reference's numbers
    must come out of the same methods the GPU tests drive."""
    o1, o2, on = t(g["o1"]), t(g["o2"]), t(g["on"])
    m = object.__new__(unsupervised_models.UnsupervisedGraphsage)
    m.link_pred_layer = BipartiteEdgePredLayer(16, 16, None, device="cpu")
    m.weight_decay, m.aggregators = 0.0, []
    m._passes = lambda b1, b2: (o1, o2, on, None)
    loss = m.loss(None, None)
    assert close(float(loss), float(g["loss_xent"]) / o1.shape[0])              # models.py:378: loss / batch_size
    assert close(float(m.mrr()), float(g["mrr"]))

    s = object.__new__(supervised_models.SupervisedGraphsage)
    agg = lambda **v: type("A", (), {"vars": v})()                              # noqa: E731
    s.aggregators = [agg(neigh_weights=t(g["a0_nw"]), self_weights=t(g["a0_sw"])), agg(weights=t(g["a1_w"]))]
    s.node_pred_vars = {"weights": t(g["head_w"]), "bias": t(g["head_b"])}
    s.logits = lambda batch: t(g["logits"])
    for sig, labels, tag in ((True, g["multi"], "sig"), (False, g["onehot"], "soft")):
        s.sigmoid_loss = sig
        for wd in (0.0, 0.05):
            s.weight_decay = wd
            assert close(float(s.loss(None, t(labels))), float(g["sup_%s_wd%d" % (tag, int(wd > 0))])), (tag, wd)
        assert close(s.predict(None).numpy(), g["pred_" + tag])
