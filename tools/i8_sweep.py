"""int8 vs DMMA (or generic, d > 64) score kernel over the dimension: where is the crossover?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysot_b200 as pb
from pysot_b200 import _lib
n, m = 20000, 1 << 19
for kernel in ("cubic", "tps"):
    for d in (10, 20, 32, 36, 40, 48, 64, 96, 128):
        rng = np.random.default_rng(0)
        X = rng.random((n, d))
        s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d), kernel=pb.TPSKernel() if kernel == "tps" else pb.CubicKernel(), tail=pb.LinearTail(d))
        s.add_points(X, np.sin(3 * X.sum(1)))
        s.set_coefficients(rng.standard_normal(n + d + 1))
        mm = m if d <= 64 else m // 4
        cand = torch.rand((mm, d), dtype=torch.float64, device="cuda")
        out = {}
        for mode in (0, 1):
            s.handle.set_option(_lib.OPT_I8_DIST, mode)
            for _ in range(2):
                idx, _ = pb.merit_select_indices(1, s, X, cand, [0.5], None)
            out[mode] = (s.handle.last_score_ms * m / mm, idx.tolist())
        print("%s d=%3d: other kernel %.2f ms, int8 %.2f ms per 2^19 x 20000 (ratio %.2f) picks %s" % (kernel, d, out[0][0], out[1][0], out[0][0] / out[1][0], "same" if out[0][1] == out[1][1] else "DIFF"), flush=True)
        del s, cand
