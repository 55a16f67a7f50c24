"""Developer check: tensor-core kernels (K3 GEMM, K4 pooled MLP) - parity tests, then timings for both MMA issue forms
(tuning mma_issue = 1: warp-uniform elect.sync issue, 0: one thread inside `if (lane == 0)`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import numpy as np  # noqa: E402
import pytest  # noqa: E402
import torch  # noqa: E402

rc = 0 if os.environ.get("TC_CHECK_SKIP_TESTS", "0") == "1" else pytest.main(["tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-k",
                  "aggregators_golden or khop_golden or sage_gemm_math or maxpool or meanpool or small_layer or full_size_forward"])
print("PYTEST_RC", int(rc), flush=True)

import graphsage_b200 as gs  # noqa: E402
from graphsage_b200 import ops  # noqa: E402

lib = gs._lib.lib()
dev = torch.device("cuda")


def timeit(fn, n=40):
    for _ in range(5):
        fn()
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3


M, F = 5632, 602
P = ops.pad_cols(F)
xs, xm = torch.randn((M, P), device=dev), torch.randn((M, P), device=dev)
Ws, Wn = torch.randn(F, 128, device=dev), torch.randn(F, 128, device=dev)
B, H = 512, 512
n_rows = 232966
table = torch.randn((n_rows, P), device=dev).to(torch.bfloat16)
ids = torch.randint(0, n_rows - 1, (B * 10 * 25,), device=dev, dtype=torch.int32)
Wm = torch.randn(F, H, device=dev) / 25.0
bm = torch.randn(H, device=dev) * 0.1
pk = ops.PackedMlpWeights()
for flag in (1, 0):
    lib.gs_set_tuning(b"mma_issue", flag)
    for math in ("tf32x3", "bf16"):
        packed = ops.PackedWeights()
        code = gs.aggregators._MATH_NAMES[math]
        t = timeit(lambda: ops.sage_gemm([(xs, F, Ws), (xm, F, Wn)], combine=ops.COMBINE_CONCAT, act=ops.ACT_RELU, math=code,
                                         packed=packed))
        print("mma_issue=%d  K3 gemm [5632 x (602+602)] -> 256 %s: %.1f us" % (flag, math, t), flush=True)
    t = timeit(lambda: ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, Wm, bm, pk, row_ids=ids), n=20)
    print("mma_issue=%d  K4 maxpool hop2 (B=512, hidden 512): %.1f us  %.1f TFLOP/s" % (flag, t, 2.0 * B * 250 * F * H / t / 1e6), flush=True)
lib.gs_set_tuning(b"mma_issue", 1)
for ksel in (0, 1):
    lib.gs_set_tuning(b"k4_kernel", ksel)
    t = timeit(lambda: ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, Wm, bm, pk, row_ids=ids), n=20)
    print("k4_kernel=%d (0 wide / 1 round-1)  K4 maxpool hop2: %.1f us  %.1f TFLOP/s" % (ksel, t, 2.0 * B * 250 * F * H / t / 1e6), flush=True)
    ids1 = ids[:B * 10]
    t = timeit(lambda: ops.maxpool_mlp_fused(table[:, :F], B, 10, Wm, bm, pk, row_ids=ids1), n=20)
    print("k4_kernel=%d  K4 maxpool hop1 (512 groups of 10): %.1f us" % (ksel, t), flush=True)
lib.gs_set_tuning(b"k4_kernel", 1)

if os.environ.get("TC_CHECK_G4", "0") == "1":
    # experimental K4 producers (1: TMA gather4, 2: + cluster multicast; hidden 128 has one slice, so 2 falls back to 1):
    # same inputs, all producers, outputs must be identical bit for bit
    # (same MMAs in the same order on the same operand bits)
    for nb, kk, KK, HH in ((3, 25, 602, 512), (77, 10, 602, 512), (5, 25, 50, 128), (512 * 10, 25, 602, 512)):
        tb = torch.randn((4096 if nb < 1000 else n_rows, ops.pad_cols(KK)), device=dev).to(torch.bfloat16)
        rid = torch.randint(0, tb.shape[0], (nb * kk,), device=dev, dtype=torch.int32)
        W = torch.randn(KK, HH, device=dev) / 25.0
        b = torch.randn(HH, device=dev) * 0.1
        outs = []
        for flag in (0, 1, 2, -1):
            lib.gs_set_tuning(b"k4_producer", flag)
            outs.append(ops.maxpool_mlp_fused(tb[:, :KK], nb, kk, W, b, ops.PackedMlpWeights(), row_ids=rid).clone())
            torch.cuda.synchronize()
        print("k4_producer 1 / 2 / -1 vs 0  groups=%d k=%d K=%d hidden=%d: max |diff| = %.3g / %.3g / %.3g" % (
            nb, kk, KK, HH, float((outs[0] - outs[1]).abs().max()), float((outs[0] - outs[2]).abs().max()),
            float((outs[0] - outs[3]).abs().max())), flush=True)
    for flag in (0, 1, 2, -1):
        lib.gs_set_tuning(b"k4_producer", flag)
        t = timeit(lambda: ops.maxpool_mlp_fused(table[:, :F], B * 10, 25, Wm, bm, pk, row_ids=ids), n=20)
        print("k4_producer=%d  K4 maxpool hop2: %.1f us  %.1f TFLOP/s" % (flag, t, 2.0 * B * 250 * F * H / t / 1e6), flush=True)
    lib.gs_set_tuning(b"k4_producer", 0)
lib.gs_set_tuning(b"k4_kernel", 0)
