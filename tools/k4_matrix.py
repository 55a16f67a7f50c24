"""K4 timing matrix at the Reddit hop-2 shape: kernel geometry x timing probes (k4_dbg: 1 zero-fill, 2 no copies)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphsage_b200 as gs
from graphsage_b200 import ops
lib = gs._lib.lib()
dev = torch.device("cuda")
N, F, H, B = 232965, 602, 512, 512
table = torch.zeros((N + 1, ops.pad_cols(F)), dtype=torch.bfloat16, device=dev)
table[:N, :F] = torch.randn((N, F), device=dev).to(torch.bfloat16)
W = torch.randn(F, H, device=dev) / 25.0
bias = torch.randn(H, device=dev)
packed = ops.PackedMlpWeights()
rs = np.random.RandomState(0)
sets = [torch.from_numpy(rs.randint(0, N, size=B * 250).astype(np.int32)).to(dev) for _ in range(8)]


def timeit(n=16):
    for i in range(3):
        ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, W, bias, packed, row_ids=sets[i])
    evs = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, W, bias, packed, row_ids=sets[i % 8]); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3


flops = 2.0 * B * 250 * F * H
for name, ksel, prod, tile, pipes, cluster in (("tmem 128 + cluster of 4", 0, 1, 128, 1, 4), ("tmem 128 + cluster of 2", 0, 1, 128, 1, 2), ("tmem 128-row tiles", 0, 1, 128, 1, 0),
                                               ("tmem 128 x 2 pipelines", 0, 1, 128, 2, 0), ("tmem 256-row tiles", 0, 1, 256, 1, 0),
                                               ("wide128 + TMA gather4", 3, 1, 128, 1, 0), ("wide128 + cp.async", 3, 0, 128, 1, 0),
                                               ("wide256 + cp.async", 2, 0, 256, 1, 0), ("round 1", 1, 0, 128, 1, 0)):
    if os.environ.get("K4_MATRIX_ONLY") and not name.startswith(os.environ["K4_MATRIX_ONLY"]):
        continue
    lib.gs_set_tuning(b"k4_kernel", ksel)
    lib.gs_set_tuning(b"k4_wide_producer", prod)
    lib.gs_set_tuning(b"k4_tile", tile)
    lib.gs_set_tuning(b"k4_pipes", pipes)
    lib.gs_set_tuning(b"k4_cluster", cluster)
    t = timeit()
    print("%-28s: %.1f us  %.1f TFLOP/s" % (name, t, flops / t / 1e6), flush=True)
lib.gs_set_tuning(b"k4_kernel", 0)
lib.gs_set_tuning(b"k4_wide_producer", 1)
lib.gs_set_tuning(b"k4_tile", 128)
lib.gs_set_tuning(b"k4_pipes", 1)
lib.gs_set_tuning(b"k4_cluster", 2)
