"""A few strategy-sized proposals (extension + device-drawn candidates + score + select), d = 30, n ~ 1000: target of
the ncu launch list behind the per-proposal breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pysot_b200 as pb
from conftest import Problem
d, n = 30, 1000
rng = np.random.default_rng(0)
prob = Problem(-5 * np.ones(d), 5 * np.ones(d))
X = prob.lb + (prob.ub - prob.lb) * rng.random((n + 40, d))
f = np.sin(X.sum(1))
pb.set_candidate_generator("device", seed=1)
s = pb.GpuRBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub)
s.add_points(X[:n], f[:n]); s._fit()
for i in range(n, n + 30):
    s.add_points(X[i], float(f[i]))
    pb.candidate_srbf(num_pts=1, opt_prob=prob, surrogate=s, X=s.X, fX=s.fX, weights=[0.5], Xpend=X[n + 30:n + 37], sampling_radius=0.2, num_cand=100 * d)
print("ok", s.fit_kind, s.handle.last_score_kernel)
