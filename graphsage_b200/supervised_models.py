"""SupervisedGraphsage - the training step around the hot path (reference
graphsage/supervised_models.py:10-126).

Forward: the B200 kernels (sample -> fused gather+mean -> tcgen05 / fp32 GEMM), wrapped in
torch.autograd.Function so the step is differentiable.  Backward: the gradient formulas of the mean / GCN
aggregators, with the weight-gradient GEMMs (X^T dZ) as plain library matmuls (torch / cuBLAS fp32) - features
are not trainable (identity_dim = 0), so nothing is scattered into the table.  Head (l2_normalize -> Dense ->
sigmoid / softmax cross-entropy + weight decay), gradient clipping to +-5 and Adam follow
supervised_models.py:85-126.
"""
import torch

from . import ops
from .layers import act_code, identity, relu  # noqa: F401
from .models import SampleAndAggregate


class _AggregateRowsFn(torch.autograd.Function):
    """y = agg.aggregate_rows(src, segments) for MeanAggregator / GCNAggregator, differentiable w.r.t. the
    aggregator weights and (for layers >= 1, where rows are addressed by ranges) w.r.t. src."""

    @staticmethod
    def forward(ctx, agg, src, segments, *weights):
        kind = "gcn" if "weights" in agg.vars else "mean"
        code, post = act_code(agg.act)
        if post is not None:
            raise NotImplementedError("training supports act=relu or identity")
        with torch.no_grad():
            if kind == "mean":
                xs, xm = ops.gather_mean(src, segments, want_self=True)
            else:
                xs, xm = None, ops.gather_mean(src, segments, include_self=True, want_self=False)[1]
            F_in = src.shape[1]
            if kind == "mean":
                parts = [(xs, F_in, weights[0]), (xm, F_in, weights[1])]
                combine = ops.COMBINE_CONCAT if agg.concat else ops.COMBINE_ADD
            else:
                parts, combine = [(xm, F_in, weights[0])], ops.COMBINE_ADD
            y = ops.sage_gemm(parts, combine=combine, bias=agg.vars.get("bias"), act=code, math=agg.math)
        ctx.kind, ctx.relu, ctx.concat = kind, code == ops.ACT_RELU, bool(agg.concat)
        ctx.segments, ctx.src_shape, ctx.F_in = segments, tuple(src.shape), F_in
        ctx.src_needs_grad = bool(torch.is_tensor(src) and src.requires_grad)
        ctx.has_bias = "bias" in agg.vars
        ctx.save_for_backward(xm if xs is None else xs, xm, y, *weights)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, xm, y = ctx.saved_tensors[:3]
        weights = ctx.saved_tensors[3:]
        F_in = ctx.F_in
        dz = dy * (y > 0).to(dy.dtype) if ctx.relu else dy
        grads_w, dsrc = [], None
        if ctx.kind == "mean":
            Ws, Wn = weights
            D = Ws.shape[1]
            dz_s, dz_n = (dz[:, :D], dz[:, D:]) if ctx.concat else (dz, dz)
            grads_w = [xs[:, :F_in].t() @ dz_s, xm[:, :F_in].t() @ dz_n]         # dW = X^T dZ  (library GEMM)
            if ctx.src_needs_grad:
                dxs, dxm = dz_s @ Ws.t(), dz_n @ Wn.t()
        else:
            (W,) = weights
            grads_w = [xm[:, :F_in].t() @ dz]
            if ctx.src_needs_grad:
                dxm = dz @ W.t()
                dxs = None
        if ctx.src_needs_grad:
            dsrc = torch.zeros(ctx.src_shape, dtype=dy.dtype, device=dy.device)
            for s in ctx.segments:
                if s.self_ids is not None or s.neigh_ids is not None:
                    raise NotImplementedError("gradient w.r.t. an id-addressed source (trainable features) is out of scope")
                n, k = s.n, s.k
                rows = slice(s.out_row0, s.out_row0 + n)
                div = float(k + (1 if ctx.kind == "gcn" else 0))
                dsrc[s.neigh_row0:s.neigh_row0 + n * k].view(n, k, -1).add_((dxm[rows] / div).unsqueeze(1))
                if ctx.kind == "gcn":
                    dsrc[s.self_row0:s.self_row0 + n].add_(dxm[rows] / div)
                else:
                    dsrc[s.self_row0:s.self_row0 + n].add_(dxs[rows])
        return (None, dsrc, None) + tuple(grads_w)


def differentiable_outputs(model, batch, normalize=True):
    """sample -> aggregate (-> l2_normalize) with an autograd graph over the aggregator weights; `model` is a
    SampleAndAggregate whose .aggregators exist (reference models.py:347-350 / supervised_models.py:79-85)."""
    batch = batch.to(device=model.device, dtype=torch.int32).reshape(-1)
    n = batch.numel()
    with torch.no_grad():
        samples, support = model.sample(batch, model.layer_infos, batch_size=n)
    num_samples = [info.num_samples for info in model.layer_infos]
    L = len(num_samples)
    counts = [n * support[h] for h in range(L + 1)]
    src = model.features
    for layer in range(L):
        hops = L - layer
        row0 = [sum(counts[:h]) for h in range(hops + 1)]
        segs = []
        for hop in range(hops):
            k = num_samples[L - hop - 1]
            if layer == 0:
                segs.append(ops.Seg(counts[hop], k, self_ids=samples[hop], neigh_ids=samples[hop + 1],
                                    out_row0=row0[hop]))
            else:
                segs.append(ops.Seg(counts[hop], k, self_row0=row0[hop], neigh_row0=row0[hop + 1],
                                    out_row0=row0[hop]))
        agg = model.aggregators[layer]
        ws = (agg.vars["weights"],) if "weights" in agg.vars else (agg.vars["self_weights"], agg.vars["neigh_weights"])
        src = _AggregateRowsFn.apply(agg, src, segs, *ws)
    out = src[:counts[0]]
    if normalize:
        out = out / torch.sqrt(torch.clamp((out * out).sum(dim=1, keepdim=True), min=1e-12))   # tf.nn.l2_normalize
    return out


def build_aggregators(model):
    """One aggregator per layer, as SampleAndAggregate.aggregate creates them (reference models.py:303-315)."""
    L = len(model.layer_infos)
    aggs = []
    for layer in range(L):
        dim_mult = 2 if model.concat and layer != 0 else 1
        act = identity if layer == L - 1 else relu
        aggs.append(model.aggregator_cls(dim_mult * model.dims[layer], model.dims[layer + 1], act=act, dropout=0.,
                                         concat=model.concat, device=model.device))
    return aggs


def classification_loss(logits, labels, sigmoid_loss):
    """reference supervised_models.py:109-117: mean over ALL elements of the sigmoid cross-entropy (multi-label), or the
    mean over nodes of the softmax cross-entropy."""
    if sigmoid_loss:
        return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="mean")
    return (-(labels * torch.log_softmax(logits, dim=1)).sum(dim=1)).mean()


def weight_decay_term(params, weight_decay):
    """weight_decay * tf.nn.l2_loss(var) = weight_decay * sum(var^2) / 2 over every variable (supervised_models.py:103-107)."""
    total = None
    for p in params:
        t = weight_decay * 0.5 * (p * p).sum()
        total = t if total is None else total + t
    return total


class SupervisedGraphsage(SampleAndAggregate):
    """Supervised GraphSAGE (reference graphsage/supervised_models.py:10-126): the hot path, then
    l2_normalize -> Dense(-> num_classes) -> sigmoid / softmax cross-entropy (+ weight decay), gradients clipped to
    [-5, 5], Adam.  TF FLAGS become constructor arguments (learning_rate, weight_decay)."""

    def __init__(self, num_classes, placeholders, features, adj, degrees, layer_infos, concat=True,
                 aggregator_type="mean", model_size="small", sigmoid_loss=False, identity_dim=0, learning_rate=0.01,
                 weight_decay=0.0, device="cuda", **kwargs):
        super(SupervisedGraphsage, self).__init__(placeholders, features, adj, degrees, layer_infos, concat=concat,
                                                  aggregator_type=aggregator_type, model_size=model_size,
                                                  identity_dim=identity_dim, device=device, **kwargs)
        if aggregator_type not in ("mean", "gcn"):
            raise NotImplementedError("training is implemented for the mean and gcn aggregators")
        self.num_classes = num_classes
        self.sigmoid_loss = sigmoid_loss
        self.learning_rate, self.weight_decay = learning_rate, weight_decay
        self.build()

    def build(self):
        from .inits import glorot, zeros
        self.aggregators = build_aggregators(self)
        dim_mult = 2 if self.concat else 1
        self.node_pred_vars = {"weights": glorot([dim_mult * self.dims[-1], self.num_classes], device=self.device),
                               "bias": zeros([self.num_classes], device=self.device)}   # supervised_models.py:88-90
        for p in self.parameters():
            p.requires_grad_(True)
        self.optimizer = torch.optim.Adam(self.parameters(), lr=self.learning_rate)      # TF AdamOptimizer defaults

    def parameters(self):
        ps = []
        for a in self.aggregators:
            ps.extend(a.vars.values())
        ps.extend(self.node_pred_vars.values())
        return ps

    def outputs(self, batch):
        """l2-normalised node representations, differentiable (supervised_models.py:79-85)."""
        return differentiable_outputs(self, batch)

    def logits(self, batch):
        return self.outputs(batch) @ self.node_pred_vars["weights"] + self.node_pred_vars["bias"]

    def loss(self, batch, labels):
        """supervised_models.py:101-118: weight decay * l2_loss(var) over aggregator + head variables, then the
        mean of the per-element sigmoid xent (multi-label) or the mean of the per-node softmax xent."""
        logits = self.logits(batch)
        labels = labels.to(device=logits.device, dtype=torch.float32)
        loss = classification_loss(logits, labels, self.sigmoid_loss)
        if self.weight_decay:
            loss = loss + weight_decay_term(self.parameters(), self.weight_decay)
        return loss

    def train_step(self, batch, labels):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss(batch, labels)
        loss.backward()
        for p in self.parameters():                                              # clip_by_value(grad, -5, 5)  :93-94
            if p.grad is not None:
                p.grad.clamp_(-5.0, 5.0)
        self.optimizer.step()
        return loss.detach()

    def predict(self, batch):
        with torch.no_grad():
            lg = self.logits(batch)
            return torch.sigmoid(lg) if self.sigmoid_loss else torch.softmax(lg, dim=1)
