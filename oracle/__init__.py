"""CPU oracle for the GraphSAGE sample-and-aggregate hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``graphsage_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline - never as the product path.

Parity status (see DESIGN.md "Oracle"):
  * The reference (williamleif/GraphSAGE @ a0fdef95) ships no tests, golden
    vectors or known-answer values, and needs TensorFlow 1.x which is not
    installable here.  The restatement is therefore pinned by executing the
    reference's OWN python (graphsage/neigh_samplers.py, aggregators.py,
    layers.py, inits.py, models.py sample/aggregate, minibatch.py construct_adj)
    under a numpy shim of the handful of TF ops it calls
    (tests/golden/tf_shim.py, tests/golden/make_golden.py) and committing the
    resulting vectors under tests/golden/*.npz.
  * What that pins: op composition, reshape/nesting order, concat order,
    divisor conventions, dummy-row behaviour, weight shapes and names.
  * What stays "parity unpinned": TensorFlow's RNG stream for
    tf.random_shuffle (the contract here is Philox4x32-10, documented in
    oracle/philox.py) and Eigen's fp32 summation order (tolerance 1e-4 rel).
"""
from .philox import philox4x32_10, mulhi32
from .sampler import (perm_prefix, sample_padded, sample_csr, sample_unigram,
                      STREAM_PADDED, STREAM_CSR)
from .aggregate import (gather_rows, mean_aggregator, gcn_aggregator,
                        maxpool_aggregator, meanpool_aggregator, dense, l2_normalize, glorot_range,
                        sample_khop, aggregate_khop, forward_2hop)
from .adjacency import construct_adj, construct_test_adj, build_padded_adj
from . import rmat

__all__ = [n for n in dir() if not n.startswith("_")]
