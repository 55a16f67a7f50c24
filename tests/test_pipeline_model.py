"""The K4 stage hand-off protocol (tools/pipeline_model.py): slot / parity arithmetic of the default kernel and of the
two TMA gather4 variants, under a random scheduler.  CPU only; guards the arithmetic the kernels share with the model."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

import pipeline_model as pm  # noqa: E402


def test_handoff_protocols_are_live_and_alias_free():
    assert pm.sweep(seeds=2) == 2 * 5 * 2 * 4


def test_model_rejects_more_stage_filling_warps_than_slots():
    msg = pm.shows_the_aliasing_bug()
    assert msg is not None and ("aliasing" in msg or "refilled" in msg or "expected" in msg)


def test_gather4_producer_tile_walk_matches_direct_indexing():
    for sa in (6, 7):
        for kblocks in (1, 2, 3, 5, 6, 7, 8, 13, 19, 20):
            for tiles in (1, 2, 5, 14):
                assert pm.check_producer_walk(sa, kblocks, tiles)


def test_round2_handoffs_are_live_and_alias_free():
    assert pm.sweep_round2(seeds=1) == 4 * (4 + 6)


def test_round2_producer_walks_match_direct_indexing():
    # the cluster kernel's producers step by MPC_PW = 4 K-blocks, the two-group cp.async producers by 2
    for step in (2, 4):
        for kblocks in (1, 2, 3, 4, 5, 9, 10, 19):
            for tiles in (1, 2, 7):
                assert pm.check_producer_walk(step, kblocks, tiles)


def test_cluster_ring_size_must_be_a_multiple_of_the_producer_warps():
    # why gs_maxpool_mlp_fused rounds n_stages down to a multiple of MPC_PW before launching the cluster kernel
    msg = pm.cluster_ring_must_be_a_multiple_of_the_producer_warps()
    assert msg is not None and ("aliasing" in msg or "refilled" in msg or "expected" in msg or "different K-blocks" in msg)
