"""One launch of the int8 score kernel at the bench shape (target of ncu).  usage: i8_one.py [log2m] [i8 0/1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysot_b200 as pb
from pysot_b200 import _lib
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d, n, m = 50, 20000, 1 << lg
rng = np.random.default_rng(0)
X = rng.random((n, d))
s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d))
s.add_points(X, np.sin(3 * X.sum(1)))
s.set_coefficients(rng.standard_normal(n + d + 1))
cand = torch.rand((m, d), dtype=torch.float64, device="cuda")
s.handle.set_option(_lib.OPT_I8_DIST, mode)
for _ in range(3):
    idx, _ = pb.merit_select_indices(1, s, X, cand, [0.5], None)
print("d=50 n=20000 m=2^%d i8=%d: %.2f ms per launch = %.3e pairs/s" % (lg, mode, s.handle.last_score_ms, m * n / s.handle.last_score_ms * 1e3))
