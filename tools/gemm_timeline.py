"""Developer probe: per-CTA timeline of the tcgen05 GEMM (globaltimer stamps of CTA (0,0))."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphsage_b200 as gs
from graphsage_b200 import ops
lib = gs._lib.lib()
dev = torch.device("cuda")
M, F = 5632, 602
P = ops.pad_cols(F)
xs = torch.randn((M, P), device=dev); xm = torch.randn((M, P), device=dev)
Ws, Wn = torch.randn(F, 128, device=dev), torch.randn(F, 128, device=dev)
names = {0: "entry", 1: "prologue done", 2: "producer: first full_a arrive", 3: "producer(g0): last arrive", 4: "epilogue wait begin",
         5: "accum ready", 6: "epilogue stores done", 7: "after final sync", 8: "mma: full_a[0] seen", 9: "mma: full_b[0] seen",
         10: "mma: it=1 ready", 11: "mma: it=8 ready", 12: "mma: last it ready", 13: "mma: last issued"}
for math in ("tf32x3", "tf32", "bf16"):
    packed = ops.PackedWeights()
    code = gs.aggregators._MATH_NAMES[math]
    for _ in range(3):
        ops.sage_gemm([(xs, F, Ws), (xm, F, Wn)], combine=ops.COMBINE_CONCAT, act=ops.ACT_RELU, math=code, packed=packed)
    buf = (ctypes.c_ulonglong * 32)()
    lib.gs_debug_read_gemm_timeline(buf, 32)
    t0 = buf[0]
    print(math, "timeline (us since CTA entry):")
    for k in sorted(names, key=lambda k: buf[k]):
        print("   %6.2f  %s" % ((buf[k] - t0) / 1e3, names[k]))
