"""Times the DMMA trailing-update kernel alone (C -= A B, K = 128) -- target of the ncu full capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysot_b200 import _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
h = _lib.Handle(0)
rng = np.random.default_rng(0)
A, B, C = rng.standard_normal((M, K)), rng.standard_normal((K, M)), rng.standard_normal((M, M))
for r in range(3):
    out, secs = h.dbg_gemm(C, A, B)
    print("gemm %dx%dx%d: %.3f ms  %.2f TFLOP/s" % (M, M, K, secs * 1e3, 2.0 * M * M * K / secs / 1e12))
err = np.max(np.abs(out[:256, :256] - (C - A @ B)[:256, :256]))
print("max err (corner)", err)
