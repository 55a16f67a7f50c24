"""Oracle for the K-hop gather, the three aggregators and the sample/aggregate
recursion.  Plain numpy, fp32 by default (pass dtype=np.float64 for the error-
budget twin).  Each function cites the reference lines it restates.

Test infrastructure - not imported by the product.
"""
import numpy as np

from .sampler import sample_padded


def relu(x):
    return np.maximum(x, 0)


def identity(x):
    return x


def glorot_range(shape):
    """reference graphsage/inits.py:15-19: U(-r, r), r = sqrt(6 / (fan_in + fan_out))."""
    return float(np.sqrt(6.0 / (shape[0] + shape[1])))


def gather_rows(features, ids):
    """tf.nn.embedding_lookup(features, ids) - reference graphsage/models.py:299."""
    return np.asarray(features)[np.asarray(ids).astype(np.int64)]


def dense(x, weights, bias=None, act=relu):
    """reference graphsage/layers.py:104-116 at dropout=0: act(x @ W + b)."""
    y = x @ weights
    if bias is not None:
        y = y + bias
    return act(y)


def mean_aggregator(self_vecs, neigh_vecs, neigh_weights, self_weights, concat=False, act=relu):
    """reference graphsage/aggregators.py:43-64 (dropout=0, bias dead - SURVEY appendix A)."""
    neigh_means = neigh_vecs.mean(axis=1, dtype=neigh_vecs.dtype)          # :48
    from_neighs = neigh_means @ neigh_weights                               # :51
    from_self = self_vecs @ self_weights                                    # :53
    out = np.concatenate([from_self, from_neighs], axis=1) if concat else from_self + from_neighs  # :55-58
    return act(out)                                                         # :64


def gcn_aggregator(self_vecs, neigh_vecs, weights, act=relu):
    """reference graphsage/aggregators.py:101-116: mean over neighbours AND self, one weight."""
    allv = np.concatenate([neigh_vecs, self_vecs[:, None, :]], axis=1)      # :106-107
    means = allv.mean(axis=1, dtype=allv.dtype)
    return act(means @ weights)                                             # :110-116


def maxpool_aggregator(self_vecs, neigh_vecs, mlp_weights, mlp_bias, neigh_weights, self_weights,
                       concat=False, act=relu):
    """reference graphsage/aggregators.py:168-195 with Dense(relu, bias) graphsage/layers.py:104-116."""
    n, k, d = neigh_vecs.shape
    h = dense(neigh_vecs.reshape(n * k, d), mlp_weights, mlp_bias, relu)   # :176-180
    h = h.reshape(n, k, -1).max(axis=1)                                     # :181-182
    from_neighs = h @ neigh_weights                                         # :184
    from_self = self_vecs @ self_weights                                    # :185
    out = np.concatenate([from_self, from_neighs], axis=1) if concat else from_self + from_neighs
    return act(out)


def meanpool_aggregator(self_vecs, neigh_vecs, mlp_weights, mlp_bias, neigh_weights, self_weights,
                        concat=False, act=relu):
    """reference graphsage/aggregators.py:246-273: as max-pool with reduce_mean over the fanout."""
    n, k, d = neigh_vecs.shape
    h = dense(neigh_vecs.reshape(n * k, d), mlp_weights, mlp_bias, relu)
    h = h.reshape(n, k, -1).mean(axis=1, dtype=h.dtype)                     # :262
    from_neighs = h @ neigh_weights
    from_self = self_vecs @ self_weights
    out = np.concatenate([from_self, from_neighs], axis=1) if concat else from_self + from_neighs
    return act(out)


def l2_normalize(x, eps=1e-12):
    """tf.nn.l2_normalize(x, 1): x * rsqrt(max(sum(x^2), eps)) - reference graphsage/models.py:368."""
    ss = (x * x).sum(axis=1, keepdims=True, dtype=x.dtype)
    return x / np.sqrt(np.maximum(ss, eps))


def sample_khop(adj, seeds, num_samples, seed, counter0):
    """reference graphsage/models.py:254-275.  num_samples is in LAYER order
    ([25, 10] = samples_1, samples_2); hop k uses num_samples[L-k-1].  Call number
    k of the recursion uses RNG counter counter0 + k."""
    samples = [np.asarray(seeds).astype(np.int32)]
    support = 1
    support_sizes = [1]
    L = len(num_samples)
    for k in range(L):
        t = L - k - 1
        support *= num_samples[t]
        node = sample_padded(adj, samples[k], num_samples[t], seed, counter0 + k)
        samples.append(node.reshape(-1))
        support_sizes.append(support)
    return samples, support_sizes


def _apply(agg, self_vecs, neigh_vecs, concat, act):
    kind = agg["type"]
    if kind == "mean":
        return mean_aggregator(self_vecs, neigh_vecs, agg["neigh_weights"], agg["self_weights"], concat, act)
    if kind == "gcn":
        return gcn_aggregator(self_vecs, neigh_vecs, agg["weights"], act)
    if kind == "maxpool":
        return maxpool_aggregator(self_vecs, neigh_vecs, agg["mlp_weights"], agg["mlp_bias"],
                                  agg["neigh_weights"], agg["self_weights"], concat, act)
    raise ValueError(kind)


def aggregate_khop(samples, features, num_samples, support_sizes, batch_size, aggregators, concat):
    """reference graphsage/models.py:278-330.  `aggregators` is a list (one per
    layer) of dicts {"type", weights...}; last layer has identity activation (:307-310)."""
    hidden = [gather_rows(features, s) for s in samples]                    # :299
    L = len(num_samples)
    for layer in range(L):
        act = identity if layer == L - 1 else relu
        nxt = []
        for hop in range(L - layer):                                        # :321
            d = hidden[hop + 1].shape[1]
            neigh = hidden[hop + 1].reshape(batch_size * support_sizes[hop], num_samples[L - hop - 1], d)  # :323-327
            nxt.append(_apply(aggregators[layer], hidden[hop], neigh, concat, act))
        hidden = nxt
    return hidden[0]


def forward_2hop(adj, features, seeds, num_samples, aggregators, concat, seed, counter0, normalize=False):
    samples, support = sample_khop(adj, seeds, num_samples, seed, counter0)
    out = aggregate_khop(samples, features, num_samples, support, len(seeds), aggregators, concat)
    return l2_normalize(out) if normalize else out
