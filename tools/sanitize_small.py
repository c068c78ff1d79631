"""Small calls through every score / gradient / fit path, meant to run under compute-sanitizer (memcheck, racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
from pysot_b200 import _lib

rng = np.random.default_rng(0)
for d, n in ((10, 300), (20, 700), (50, 257)):
    lb, ub = np.zeros(d), np.ones(d)
    cen = rng.random((4, d))
    X = np.clip(cen[rng.integers(0, 4, n)] + 10.0 ** rng.uniform(-5, -1, (n, 1)) * rng.standard_normal((n, d)), 0, 1)
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
    s.add_points(X, np.sin(X.sum(1)))
    s._fit()
    for k in range(3):  # extensions
        x = np.clip(cen[0] + 1e-2 * rng.standard_normal(d), 0, 1)
        s.add_points(x, float(np.sin(x.sum())))
        s._fit()
    for m in (1, 17, 900, 7200, 15000):  # center split 4 / 2 / 1
        q = np.clip(cen[rng.integers(0, 4, m)] + 10.0 ** rng.uniform(-6, -1, (m, 1)) * rng.standard_normal((m, d)), 0, 1)
        q[: min(m, 3)] = s.X[: min(m, 3)]
        s.predict(q)
        pb.weighted_distance_merit(min(m, 3), s, s.X, s.fX, q, [0.3, 0.5, 0.95], None)
    s.predict_deriv(q[:5])
    s.predict_deriv(q[:5000])
    if d >= 28:
        s.handle.set_option(_lib.OPT_I8_DIST, 1)
        s.predict(q)
        pb.weighted_distance_merit(3, s, s.X, s.fX, q, [0.3, 0.5, 0.95], None)
    print("d=%d ok, last kernel %s" % (d, s.handle.last_score_kernel), flush=True)
