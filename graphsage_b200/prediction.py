"""Edge-prediction head applied to the hot path's outputs in the unsupervised model
(reference graphsage/prediction.py:12-128 `BipartiteEdgePredLayer`).  Plain torch on whatever device the embeddings
live on; tiny next to the hot path (a [B, D] x [D, 20] product).

The arithmetic lives in the module-level functions (pair_scores, negative_scores, *_objective, mrr_from_affinities);
`BipartiteEdgePredLayer` is the reference-shaped wrapper that owns the optional bilinear matrix."""
import torch
import torch.nn.functional as F

from .inits import glorot, zeros
from .layers import Layer

HINGE_MARGIN = 0.1                  # reference prediction.py:32


def pair_scores(u, v, bilinear=None):
    """[batch] score of each (u_i, v_i): the dot product, or u_i^T A v_i with a bilinear matrix A [dim_u, dim_v]
    (reference prediction.py:68-80)."""
    if bilinear is not None:
        v = v @ bilinear.t()
    return torch.einsum("bd,bd->b", u, v)


def negative_scores(u, negatives, bilinear=None):
    """[batch, num_neg] score of every u_i against every shared negative (reference prediction.py:82-92)."""
    if bilinear is not None:
        u = u @ bilinear
    return u @ negatives.t()


def xent_objective(pos, neg, neg_weight=1.0):
    """Sum of sigmoid cross-entropies: label 1 for the true pairs (softplus(-x)), label 0 for the negatives
    (softplus(x)), the latter scaled by `neg_weight` (reference prediction.py:102-110)."""
    return F.softplus(-pos).sum() + neg_weight * F.softplus(neg).sum()


def skipgram_objective(pos, neg):
    """sum_i (pos_i - log sum_j exp(neg_ij)), as the reference writes it (prediction.py:112-117)."""
    return (pos - torch.logsumexp(neg, dim=1)).sum()


def hinge_objective(pos, neg, margin=HINGE_MARGIN):
    """sum_ij max(0, neg_ij - pos_i + margin) (reference prediction.py:119-125)."""
    return torch.clamp_min(neg - pos[:, None] + margin, 0).sum()


def mrr_from_affinities(aff, neg_aff):
    """Mean reciprocal rank of the true pair among [negatives..., true] by descending affinity, ties to the lower
    column (reference models.py:393-405: two top_k passes = rank of the last column)."""
    table = torch.cat([neg_aff, aff[:, None]], dim=1)
    by_score = torch.argsort(table, dim=1, descending=True, stable=True)
    rank_of_column = torch.argsort(by_score, dim=1, stable=True)
    return torch.reciprocal((rank_of_column[:, -1] + 1).float()).mean()


_OBJECTIVES = ("xent", "skipgram", "hinge")


class BipartiteEdgePredLayer(Layer):
    """Skip-gram style loss between `inputs1`, their positive partners `inputs2` and negatives shared by the batch.
    loss_fn: 'xent' (default), 'skipgram', 'hinge'; `bilinear_weights` adds vars['weights'] [input_dim1, input_dim2]."""

    def __init__(self, input_dim1, input_dim2, placeholders=None, dropout=False, act=torch.sigmoid, loss_fn="xent",
                 neg_sample_weights=1.0, bias=False, bilinear_weights=False, device="cuda", **kwargs):
        if loss_fn not in _OBJECTIVES:
            raise ValueError("unknown loss_fn %r (expected one of %s)" % (loss_fn, ", ".join(_OBJECTIVES)))
        Layer.__init__(self, **kwargs)
        self.input_dim1, self.input_dim2, self.output_dim = input_dim1, input_dim2, 1
        self.act, self.bias, self.eps, self.margin = act, bias, 1e-7, HINGE_MARGIN
        self.neg_sample_weights, self.bilinear_weights, self.loss_name = neg_sample_weights, bilinear_weights, loss_fn
        self.dropout = placeholders["dropout"] if (dropout and placeholders is not None) else 0.
        if bilinear_weights:                       # xavier-uniform, as tf.contrib.layers.xavier_initializer (prediction.py:47-51)
            self.vars["weights"] = glorot([input_dim1, input_dim2], device=device)
        if bias:
            self.vars["bias"] = zeros([1], device=device)

    def _bilinear(self):
        return self.vars["weights"] if self.bilinear_weights else None

    def affinity(self, inputs1, inputs2):
        return pair_scores(inputs1, inputs2, self._bilinear())

    def neg_cost(self, inputs1, neg_samples, hard_neg_samples=None):
        return negative_scores(inputs1, neg_samples, self._bilinear())

    def loss(self, inputs1, inputs2, neg_samples):
        pos, neg = self.affinity(inputs1, inputs2), self.neg_cost(inputs1, neg_samples)
        if self.loss_name == "xent":
            return xent_objective(pos, neg, self.neg_sample_weights)
        if self.loss_name == "skipgram":
            return skipgram_objective(pos, neg)
        return hinge_objective(pos, neg, self.margin)
