"""Timing of b200rbf_extend (one appended point per call) at strategy sizes, and the target of an ncu launch list.
usage: ext_profile.py [d] [n0] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb

d = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n0 = int(sys.argv[2]) if len(sys.argv) > 2 else 900
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
rng = np.random.default_rng(0)
X = rng.random((n0 + steps, d))
f = np.sin(X.sum(1))[:, None]
s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d))
s.add_points(X[:n0], f[:n0])
t0 = time.perf_counter(); s._fit(); t_full = time.perf_counter() - t0
s.updated = False; s._fit_Xu = None
t0 = time.perf_counter(); s._fit(); t_full = time.perf_counter() - t0
full_dev = s.fit_seconds
dev, wall = [], []
l0 = s.handle.launches
for n in range(n0, n0 + steps):
    s.add_points(X[n], f[n])
    t0 = time.perf_counter()
    s._fit()
    wall.append(time.perf_counter() - t0)
    dev.append(s.fit_seconds)
    assert s.fit_kind == "extension"
print("d=%d n=%d..%d: full fit %.3f ms device / %.3f ms wall; extension per point: device median %.3f ms, wall median %.3f ms, "
      "%.1f launches per extension" % (d, n0, n0 + steps, 1e3 * full_dev, 1e3 * t_full, 1e3 * np.median(dev), 1e3 * np.median(wall),
                                       (s.handle.launches - l0) / steps))
