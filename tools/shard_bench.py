"""1-GPU microbench: the layer-0 fused gather+mean on a dense table vs the sharded resolver (one shard; with and without a
remap table) - what the partitioned address arithmetic costs when nothing is remote."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphsage_b200 as gs
from graphsage_b200 import ops, parallel
dev = torch.device("cuda")
N, F, B = 232965, 602, 512
rs = np.random.RandomState(0)
feats = rs.standard_normal((N, F)).astype(np.float32)
table = torch.zeros((N + 1, ops.pad_cols(F)), device=dev)
table[:N, :F] = torch.from_numpy(feats).to(dev)
sets = []
for i in range(6):
    s0 = torch.from_numpy(rs.randint(0, N, size=B).astype(np.int32)).to(dev)
    s1 = torch.from_numpy(rs.randint(0, N, size=B * 10).astype(np.int32)).to(dev)
    s2 = torch.from_numpy(rs.randint(0, N, size=B * 250).astype(np.int32)).to(dev)
    sets.append([ops.Seg(B, 10, self_ids=s0, neigh_ids=s1, out_row0=0), ops.Seg(B * 10, 25, self_ids=s1, neigh_ids=s2, out_row0=B)])


def timeit(src, n=24):
    for i in range(3):
        ops.gather_mean(src, sets[i])
    evs = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.gather_mean(src, sets[i % 6]); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3


print("dense table               : %.1f us" % timeit(table[:, :F]))
shard = parallel.ShardedFeatures(feats, N)
print("sharded, 1 shard, no remap: %.1f us" % timeit(shard))
remap = np.arange(N + 1, dtype=np.int32)
shard.remap = torch.from_numpy(remap).to(dev)
shard._table.remap = shard.remap.data_ptr()
print("sharded, identity remap   : %.1f us" % timeit(shard))
shard.close()
