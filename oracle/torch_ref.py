"""The reference op sequence restated 1:1 in torch-CPU (multi-threaded MKL) - the timed CPU
baseline ("port": TensorFlow 1.x cannot be installed here).

Mirrors the TF ops including the materialised [250*B, F] neighbour tensor the reference creates
(reference graphsage/models.py:299): embedding_lookup -> reduce_mean -> matmul -> concat -> relu.
Validated against oracle/aggregate.py in tests/test_oracle_golden.py.

Test/bench infrastructure - not imported by the product.
"""
import numpy as np
import torch

from .sampler import perm_prefix


def sample_padded(adj_t, ids_t, k, seed, counter):
    """neigh_samplers.py:26-28 on torch-CPU: gather whole rows, permute columns, slice."""
    rows = adj_t.index_select(0, ids_t.long())                                 # embedding_lookup   :26
    pi = torch.from_numpy(perm_prefix(seed, counter, adj_t.shape[1], adj_t.shape[1]).astype(np.int64))
    rows = rows.t().index_select(0, pi).t()                                    # transpose/shuffle/transpose :27
    return rows[:, :k].contiguous()                                            # slice :28


def forward(adj_t, feats_t, seeds_t, num_samples, aggs, concat, kind, seed, counter0, normalize=True):
    """models.py:254-330 + :368 for mean / gcn / maxpool, fp32."""
    L = len(num_samples)
    samples, support, sup = [seeds_t], [1], 1
    for k in range(L):
        t = L - k - 1
        sup *= num_samples[t]
        samples.append(sample_padded(adj_t, samples[k], num_samples[t], seed, counter0 + k).reshape(-1))
        support.append(sup)
    B = seeds_t.numel()
    hidden = [feats_t.index_select(0, s.long()) for s in samples]             # models.py:299
    for layer in range(L):
        a = aggs[layer]
        last = layer == L - 1
        nxt = []
        for hop in range(L - layer):
            d = hidden[hop + 1].shape[1]
            neigh = hidden[hop + 1].reshape(B * support[hop], num_samples[L - hop - 1], d)
            selfv = hidden[hop]
            if kind == "mean":
                fn = neigh.mean(dim=1) @ a["neigh_weights"]
                fs = selfv @ a["self_weights"]
                out = torch.cat([fs, fn], dim=1) if concat else fs + fn
            elif kind == "gcn":
                out = torch.cat([neigh, selfv[:, None, :]], dim=1).mean(dim=1) @ a["weights"]
            elif kind == "maxpool":
                n_, k_, _ = neigh.shape
                h = torch.relu(neigh.reshape(n_ * k_, d) @ a["mlp_weights"] + a["mlp_bias"])
                h = h.reshape(n_, k_, -1).max(dim=1).values
                fn = h @ a["neigh_weights"]
                fs = selfv @ a["self_weights"]
                out = torch.cat([fs, fn], dim=1) if concat else fs + fn
            else:
                raise ValueError(kind)
            nxt.append(out if last else torch.relu(out))
        hidden = nxt
    out = hidden[0]
    if normalize:
        out = out / torch.sqrt(torch.clamp((out * out).sum(dim=1, keepdim=True), min=1e-12))
    return out
