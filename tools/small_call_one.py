"""One strategy-sized scoring call (d = 20, n = 2500, 2000 candidates) and a one-point predict: targets of ncu for the
center-split launch shape of the DMMA score kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
d, n, m = 20, 2500, 2000
rng = np.random.default_rng(0)
lb, ub = -5.12 * np.ones(d), 5.12 * np.ones(d)
X = lb + (ub - lb) * rng.random((n, d))
s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
s.add_points(X, np.sin(X.sum(1)))
s.set_coefficients(rng.standard_normal(n + d + 1))
q = lb + (ub - lb) * rng.random((m, d))
for _ in range(3):
    pb.merit_select_indices(1, s, X, q, [0.5], None)
    s.predict(q[:1])
print("d=%d n=%d m=%d: score kernel %s, %.1f us per launch" % (d, n, m, s.handle.last_score_kernel, 1e3 * s.handle.last_score_ms))
