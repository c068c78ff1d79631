"""Latency of predict / score right after the model changed (extension) vs on a static model."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
d, n0 = 20, 2500
rng = np.random.default_rng(0)
lb, ub = -5.12 * np.ones(d), 5.12 * np.ones(d)
X = lb + (ub - lb) * rng.random((n0 + 300, d))
f = np.sin(X.sum(1))
s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
s.add_points(X[:n0], f[:n0]); s._fit()
h = s.handle
q1 = lb + (ub - lb) * rng.random((1, d)); q2k = lb + (ub - lb) * rng.random((2000, d))
tf, tp1, tp2, ts, tp3 = [], [], [], [], []
for i in range(n0, n0 + 200):
    s.add_points(X[i], f[i])
    t0 = time.perf_counter(); s._fit(); t1 = time.perf_counter()
    h.predict(q1, lb, ub); t2 = time.perf_counter()
    h.predict(q1, lb, ub); t3 = time.perf_counter()
    h.score(q2k, lb, ub, s.X, True); t4 = time.perf_counter()
    h.predict(q1, lb, ub); t5 = time.perf_counter()
    tf.append(t1 - t0); tp1.append(t2 - t1); tp2.append(t3 - t2); ts.append(t4 - t3); tp3.append(t5 - t4)
med = lambda v: 1e3 * float(np.median(v))
print("fit(kind=%s) %.3f ms | first predict after it %.3f | second predict %.3f | score %.3f | predict after score %.3f"
      % (s.fit_kind, med(tf), med(tp1), med(tp2), med(ts), med(tp3)))
