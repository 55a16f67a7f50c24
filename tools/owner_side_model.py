#!/usr/bin/env python
"""CPU model of the NVLink row traffic of the node-partitioned gather on the bench graph, now and with owner-side
aggregation of remote hop-1 nodes (DESIGN 'What is left' item 1).  Draws follow the padded-table sampler's distribution
(one column permutation per hop); every rank's replica set is the one bench.py installs (parallel.hot_remote_rows).

owner-side aggregation: a hop-1 node u that is neither owned nor replicated here is served by its owner, which reads u's
25 sampled neighbours through ITS own table (own rows + its replicas local, the rest over NVLink) and ships two rows back
(u's self row and the neighbour mean).  Exact for the mean / GCN aggregators: the column permutation is shared per call."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from graphsage_b200 import parallel as par  # noqa: E402


def main():
    g = bench.build_graph()
    adj, N = g["adj"], bench.N_NODES
    rng = np.random.RandomState(0)
    for world in (2, 4, 8):
        rs = np.array(par.community_bounds(g["comm"], world))
        for frac, label in ((0.0, "none"), (0.125, "1/8"), (0.25, "1/4")):
            cache = int(N * frac)
            local = np.zeros((world, N + 1), bool)
            for r in range(world):
                local[r, rs[r]:rs[r + 1]] = True
                local[r, N] = True
                if cache:
                    local[r, par.hot_remote_rows(adj, N, world, r, cache, list(rs))] = True
            now = owner = rows = 0
            for rank in range(world):
                B = 512 * 4
                seeds = rng.randint(rs[rank], rs[rank + 1], size=B)
                h1 = adj[seeds][:, rng.permutation(adj.shape[1])[:10]].reshape(-1)
                h2 = adj[h1][:, rng.permutation(adj.shape[1])[:25]]
                rem1, rem2 = ~local[rank][h1], ~local[rank][h2]
                now += rem1.sum() + rem2.sum()
                o1 = np.searchsorted(rs[1:-1], np.minimum(h1, N - 1), side="right")
                at_owner = ~local[o1[:, None], h2]                       # rows the OWNER of u would pull over NVLink
                owner += 2 * rem1.sum() + at_owner[rem1].sum() + rem2[~rem1].sum()
                rows += B * 261
            print("world %d  replicas %-4s  remote rows: now %.1f %%   owner-side aggregation %.1f %%" % (
                world, label, 100.0 * now / rows, 100.0 * owner / rows), flush=True)


if __name__ == "__main__":
    main()
