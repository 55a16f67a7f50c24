"""UniformNeighborSampler - the surface of reference graphsage/neigh_samplers.py:15-29 over the
K1 sampler kernels."""
import torch

from . import ops
from .layers import Layer


class UniformNeighborSampler(Layer):
    """Uniformly samples neighbours.  Assumes adj lists are padded with random re-sampling
    (reference graphsage/minibatch.py:227-245).

    sampler((ids, num_samples)) -> int32 [n, num_samples] with out[i, j] = adj_info[ids[i], pi[j]] for ONE
    column permutation pi per call (neigh_samplers.py:26-28).  pi comes from Philox4x32-10 keyed by
    `seed`, counter = number of calls so far (oracle/sampler.py documents the contract).
    """

    def __init__(self, adj_info, seed=123, **kwargs):
        super(UniformNeighborSampler, self).__init__(**kwargs)
        self.set_adj(adj_info)
        self.seed = int(seed)
        self.counter = 0          # one fresh permutation per call, like a new random_shuffle op execution
        self.counter_dev = None   # optional device uint64 added to `counter` (CUDA-graph replay)

    def set_adj(self, adj_info):
        """Swap the table (train adj <-> test adj), as tf.assign does in supervised_train.py:260-261."""
        if adj_info.dtype != torch.int32 or adj_info.dim() != 2:
            raise TypeError("adj_info must be an int32 [N+1, max_degree] tensor")
        ops.require_cuda(adj_info)
        self.adj_info = adj_info.contiguous()

    def _call(self, inputs):
        ids, num_samples = inputs
        out = ops.sample_padded(self.adj_info, ids, int(num_samples), self.seed, self.counter,
                                counter_dev=self.counter_dev)
        self.counter += 1
        return out

    def sample_khop(self, ids, fanouts):
        """All hops of SampleAndAggregate.sample (reference models.py:254-275) in one kernel launch;
        identical to len(fanouts) successive calls.  fanouts in hop order, e.g. [10, 25]."""
        outs = ops.sample_padded_khop(self.adj_info, ids, [int(k) for k in fanouts], self.seed, self.counter,
                                      counter_dev=self.counter_dev)
        self.counter += len(fanouts)
        return outs


class CSRNeighborSampler(Layer):
    """Per-node uniform draws from a CSR adjacency (north_star's warp-per-node mode; the reference
    has only the padded table).  Same call convention as UniformNeighborSampler."""

    def __init__(self, indptr, indices, seed=123, replace_if_short=True, pad_id=None, **kwargs):
        super(CSRNeighborSampler, self).__init__(**kwargs)
        ops.require_cuda(indptr, indices)
        self.indptr, self.indices = indptr.contiguous(), indices.contiguous()
        self.seed = int(seed)
        self.counter = 0
        self.counter_dev = None
        self.replace_if_short = replace_if_short
        self.pad_id = int(indptr.numel() - 1) if pad_id is None else int(pad_id)   # dummy node N

    def _call(self, inputs):
        ids, num_samples = inputs
        out = ops.sample_csr(self.indptr, self.indices, ids, int(num_samples), self.seed, self.counter,
                             self.replace_if_short, self.pad_id, counter_dev=self.counter_dev)
        self.counter += 1
        return out
