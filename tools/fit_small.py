"""Fit latency at strategy-scale sizes, LU vs Cholesky (to place the automatic switch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
from pysot_b200 import _lib
for d in (10, 30):
    for n in (56, 64, 80, 100, 127, 150, 300, 500, 1000):
        if n < 2 * (d + 1) + 32: continue
        rng = np.random.default_rng(0)
        X = rng.random((n, d)); fX = np.sin(3 * X.sum(1))[:, None]
        out = []
        for method in (1, 2):
            s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d))
            s.handle.set_option(_lib.OPT_FIT_METHOD, method)
            s.add_points(X, fX); s._fit(); s._incremental = False
            ts = []
            for r in range(10):
                s.updated = False; t0 = time.perf_counter(); s._fit(); ts.append(time.perf_counter() - t0)
            out.append((min(ts) * 1e3, s.fit_seconds * 1e3, s.handle.launches // 11))
        print("d=%2d n=%5d  LU %.3f ms (dev %.3f, %d launches)   Cholesky %.3f ms (dev %.3f, %d launches)" % (d, n, *out[0], *out[1]), flush=True)
