"""Edge-prediction head applied to the hot path's outputs in the unsupervised model
(reference graphsage/prediction.py:12-128 `BipartiteEdgePredLayer`).  Plain torch on whatever device the embeddings
live on; tiny next to the hot path (a [B, D] x [D, 20] product)."""
import torch
import torch.nn.functional as F

from .inits import glorot, zeros
from .layers import Layer


class BipartiteEdgePredLayer(Layer):
    """Skip-gram style loss between `inputs1`, their positive partners `inputs2` and negatives shared by the batch.

    affinity(u, v) = u . v, or u^T A v with `bilinear_weights` (vars['weights'] [input_dim1, input_dim2]);
    loss_fn: 'xent' (default), 'skipgram', 'hinge' (reference prediction.py:102-122)."""

    def __init__(self, input_dim1, input_dim2, placeholders=None, dropout=False, act=torch.sigmoid, loss_fn="xent",
                 neg_sample_weights=1.0, bias=False, bilinear_weights=False, device="cuda", **kwargs):
        super(BipartiteEdgePredLayer, self).__init__(**kwargs)
        self.input_dim1, self.input_dim2 = input_dim1, input_dim2
        self.act, self.bias = act, bias
        self.eps = 1e-7
        self.margin = 0.1                       # hinge margin (prediction.py:32)
        self.neg_sample_weights = neg_sample_weights
        self.bilinear_weights = bilinear_weights
        self.dropout = placeholders["dropout"] if (dropout and placeholders is not None) else 0.
        self.output_dim = 1
        if bilinear_weights:
            self.vars["weights"] = glorot([input_dim1, input_dim2], device=device)    # xavier-uniform (prediction.py:47-51)
        if bias:
            self.vars["bias"] = zeros([self.output_dim], device=device)
        fns = {"xent": self._xent_loss, "skipgram": self._skipgram_loss, "hinge": self._hinge_loss}
        if loss_fn not in fns:
            raise ValueError("unknown loss_fn %r" % (loss_fn,))
        self.loss_fn = fns[loss_fn]

    def affinity(self, inputs1, inputs2):
        """[batch] scores of the pairs (prediction.py:68-80)."""
        if self.bilinear_weights:
            return (inputs1 * (inputs2 @ self.vars["weights"].t())).sum(dim=1)
        return (inputs1 * inputs2).sum(dim=1)

    def neg_cost(self, inputs1, neg_samples, hard_neg_samples=None):
        """[batch, num_neg] scores of every input against every negative (prediction.py:82-92)."""
        if self.bilinear_weights:
            inputs1 = inputs1 @ self.vars["weights"]
        return inputs1 @ neg_samples.t()

    def loss(self, inputs1, inputs2, neg_samples):
        return self.loss_fn(inputs1, inputs2, neg_samples)

    def _xent_loss(self, inputs1, inputs2, neg_samples, hard_neg_samples=None):
        # sigmoid xent with labels 1 is softplus(-x), with labels 0 softplus(x)   (prediction.py:102-110)
        aff = self.affinity(inputs1, inputs2)
        neg_aff = self.neg_cost(inputs1, neg_samples, hard_neg_samples)
        return F.softplus(-aff).sum() + self.neg_sample_weights * F.softplus(neg_aff).sum()

    def _skipgram_loss(self, inputs1, inputs2, neg_samples, hard_neg_samples=None):
        aff = self.affinity(inputs1, inputs2)
        neg_aff = self.neg_cost(inputs1, neg_samples, hard_neg_samples)
        return (aff - torch.log(torch.exp(neg_aff).sum(dim=1))).sum()                  # prediction.py:112-117 (as written there)

    def _hinge_loss(self, inputs1, inputs2, neg_samples, hard_neg_samples=None):
        aff = self.affinity(inputs1, inputs2)
        neg_aff = self.neg_cost(inputs1, neg_samples, hard_neg_samples)
        return torch.relu(neg_aff - (aff.unsqueeze(1) - self.margin)).sum()            # prediction.py:119-125


def mrr_from_affinities(aff, neg_aff):
    """Mean reciprocal rank of the true pair among [negatives..., true] by descending affinity, ties to the lower
    column (reference models.py:393-405: two top_k passes = rank of the last column)."""
    aff_all = torch.cat([neg_aff, aff.unsqueeze(1)], dim=1)
    order = torch.argsort(aff_all, dim=1, descending=True, stable=True)
    ranks = torch.argsort(order, dim=1, stable=True)
    return (1.0 / (ranks[:, -1] + 1).float()).mean()
