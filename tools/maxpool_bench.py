"""K4 micro-benchmark at the Reddit shape: hop-2 of layer 0 (128000 gathered rows x 602 -> 512, max over 25)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphsage_b200 as gs
from graphsage_b200 import ops
dev = torch.device("cuda")
N, F, H, B = 232965, 602, 512, 512
P = int(os.environ.get('MP_PITCH', ops.pad_cols(F)))
table = torch.zeros((N + 1, P), dtype=torch.bfloat16, device=dev)
table[:N, :F] = torch.randn((N, F), device=dev).to(torch.bfloat16)
W = torch.randn(F, H, device=dev) / 25.0
bias = torch.randn(H, device=dev)
packed = ops.PackedMlpWeights()
rs = np.random.RandomState(0)
sets = [torch.from_numpy(rs.randint(0, N, size=B * 250).astype(np.int32)).to(dev) for _ in range(12)]
for i in range(3):
    ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, W, bias, packed, row_ids=sets[i])
torch.cuda.synchronize()
torch.cuda._sleep(4000000)
evs = []
for i in range(3, 12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, W, bias, packed, row_ids=sets[i]); e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
t = np.median([a.elapsed_time(b) for a, b in evs])
flops = 2.0 * B * 250 * F * H
print("maxpool_mlp_fused hop2: %.1f us  %.1f TFLOP/s (algorithmic, K=602)  gathered %.1f GB/s" % (
    t * 1e3, flops / t / 1e9, B * 250 * F * 2 / t / 1e6))
import ctypes
buf = (ctypes.c_ulonglong * 128)()
gs._lib.lib().gs_debug_read_maxpool_timeline(buf, 128)
t0 = buf[0]
names = ["prod: tile start", "prod: tile issued", "mma: want acc", "mma: got acc", "mma: last issued", "epi: want acc", "epi: got acc", "epi: done"]
for tc in range(3, 7):
    print("tile %d: " % tc + "  ".join("%s=%.2f" % (names[j].split(": ")[1][:11] + "@" + names[j][:3], (buf[tc * 8 + j] - t0) / 1e3) for j in range(8)))

n = max(1, buf[104])
print("producer step cycles (avg over %d K-blocks): wait %.0f  stores %.0f  fence %.0f  arrive %.0f" % (n, buf[100]/n, buf[101]/n, buf[102]/n, buf[103]/n))

if buf[105]:          # built with GS_EXTRA_NVCC_FLAGS=-DGS_K4_EPI_PROBE
    nb = buf[105]
    print("epilogue cycles per 32-column block (avg over %d): tmem ld %.0f  staging stores %.0f  barrier %.0f  pooling + stores %.0f  "
          "barrier %.0f" % (nb, buf[100] / nb, buf[101] / nb, buf[102] / nb, buf[103] / nb, buf[104] / nb))

