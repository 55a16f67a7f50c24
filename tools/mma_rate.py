"""Developer probe: raw tcgen05.mma rate with K4's operand layout (see csrc/probe_tc.cu)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401,E402  (CUDA context)
from graphsage_b200 import _lib  # noqa: E402

lib = _lib.lib()
torch.zeros(1, device="cuda")
out = (ctypes.c_ulonglong * 2)()
for ctas in (148,):
    for ncols in (128, 256):
        for mode in (0, 4, 5, 6):
            for n in (512,):
                rc = lib.gs_debug_mma_rate(ncols, mode, n, ctas, out)
                assert rc == 0, rc
                print("ctas %3d  N=%3d  mode %d  n_mma %4d: issue loop %6.1f cyc/MMA   to completion %6.1f cyc/MMA" % (
                    ctas, ncols, mode, n, out[0] / n, out[1] / n))
