"""Times the DMMA trailing-update kernel alone (C -= A B, K = 128) -- target of the ncu full capture.
usage: gemm_profile.py M K [variant: 2 = 2-stage BK16, 3 = 3-stage BK16, 4 = 2-stage BK32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysot_b200 import _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
variants = [int(sys.argv[3])] if len(sys.argv) > 3 else [3, 2, 4]
h = _lib.Handle(0)
rng = np.random.default_rng(0)
A, B, C = rng.standard_normal((M, K)), rng.standard_normal((K, M)), rng.standard_normal((M, M))
ref = (C - A @ B)[:256, :256]
for v in variants:
    h.set_option(_lib.OPT_GEMM_STAGES, v)
    best = 1e9
    for r in range(4):
        out, secs = h.dbg_gemm(C, A, B)
        best = min(best, secs)
    print("gemm variant %d %dx%dx%d: %.3f ms  %.2f TFLOP/s  max err %.2e" % (v, M, M, K, best * 1e3, 2.0 * M * M * K / best / 1e12,
          np.max(np.abs(out[:256, :256] - ref))))
