"""Strategy-scale latencies (BASELINE configs #1/#2/#5 sizes): one proposal = add a point, refit, score 100*d
candidates, pick.  Prints wall-clock per phase on the GPU path and on the CPU oracle (same sizes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
from pysot_b200 import _lib
from oracle import rbf_oracle as orc


class Problem:
    def __init__(self, lb, ub):
        self.lb, self.ub, self.dim = lb, ub, len(lb)
        self.int_var, self.cont_var = np.array([], dtype=int), np.arange(len(lb))

def bench(d, n, reps=20):
    rng = np.random.default_rng(0)
    lb, ub = -15.0 * np.ones(d), 20.0 * np.ones(d)
    prob = Problem(lb, ub)
    X = lb + (ub - lb) * rng.random((n + reps, d))
    fX = np.sin(X.sum(1) / 10)[:, None]
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
    s.add_points(X[:n], fX[:n])
    s.predict(X[:2])
    np.random.seed(0)
    t_fit = t_gen = t_merit = 0.0
    for r in range(reps):
        s.add_points(X[n + r], fX[n + r])
        t0 = time.perf_counter(); s._fit(); t1 = time.perf_counter()
        cand = pb.generate_dycors(prob, s.X, s.fX, 0.3, num_cand=100 * d); t2 = time.perf_counter()
        pb.weighted_distance_merit(1, s, s.X, s.fX, cand, [0.5]); t3 = time.perf_counter()
        t_fit += t1 - t0; t_gen += t2 - t1; t_merit += t3 - t2
    o = orc.OracleRBF(d, lb, ub); o.add_points(X[:n], fX[:n]); o.fit()
    c_fit = c_merit = 0.0
    for r in range(min(reps, 5)):
        o.add_points(X[n + r], fX[n + r])
        t0 = time.perf_counter(); o.fit(); t1 = time.perf_counter()
        cand = orc.gen_dycors(prob, o.X, o.fX, 0.3, num_cand=100 * d)
        t2 = time.perf_counter(); orc.weighted_distance_merit(1, o, o.X, o.fX, cand, [0.5]); t3 = time.perf_counter()
        c_fit += t1 - t0; c_merit += t3 - t2
    k = min(reps, 5)
    print("d=%2d n=%5d m=%5d | GPU fit %7.3f ms (%s) merit %7.3f ms | host gen %6.3f ms | CPU oracle fit(+1 incr) %7.3f ms merit %8.3f ms"
          % (d, n, 100 * d, 1e3 * t_fit / reps, s.handle.last_fit_method, 1e3 * t_merit / reps, 1e3 * t_gen / reps,
             1e3 * c_fit / k, 1e3 * c_merit / k), flush=True)

for d, n in [(10, 30), (10, 100), (10, 300), (10, 500), (30, 200), (30, 700), (30, 1000), (30, 2000), (20, 1000), (20, 3000)]:
    bench(d, n)
