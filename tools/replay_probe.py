"""Per-call timing inside the SOP replay (device candidates): which calls are slow, as a function of n?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pysot_b200 as pb
from conftest import GOLDEN, Problem
g = dict(np.load(os.path.join(GOLDEN, "trace_sop.npz")))
d = int(g["dim"]); prob = Problem(g["lb"], g["ub"], g["int_var"])
pb.set_candidate_generator("device", seed=1)
s = pb.GpuRBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub)
pend_off = np.concatenate(([0], np.cumsum(g["pend_cnt"]))); pt_off = np.concatenate(([0], np.cumsum(g["num_pts"])))
have = 0; rows = []
for i in range(len(g["n"])):
    n = int(g["n"][i])
    t0 = time.perf_counter()
    if n > have:
        s.add_points(g["allX"][have:n], g["allfX"][have:n]); have = n
    X, fX = g["allX"][:n], g["allfX"][:n]
    t1 = time.perf_counter()
    s._fit(); t2 = time.perf_counter()
    pts = pb.candidate_dycors(num_pts=1, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=[1.0], prob_perturb=float(g["prob_perturb"][i]),
                              Xpend=g["pend"][pend_off[i]:pend_off[i + 1]], sampling_radius=float(g["sampling_radius"][i]),
                              num_cand=int(g["num_cand"][i]), xbest=g["xbest"][i].copy())
    t3 = time.perf_counter()
    ref = g["new_points"][pt_off[i]:pt_off[i + 1]]
    s.predict_deriv(ref); t4 = time.perf_counter()
    s.predict(ref); t5 = time.perf_counter()
    rows.append((n, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, s.fit_kind, s.handle.last_score_kernel))
rows_np = np.array([r[:6] for r in rows])
for lo, hi in ((0, 500), (500, 1000), (1000, 1500), (1500, 2000), (2000, 2500), (2500, 3100)):
    sel = (rows_np[:, 0] >= lo) & (rows_np[:, 0] < hi)
    if sel.any():
        md = np.median(rows_np[sel], axis=0) * 1e3; mx = rows_np[sel].max(axis=0) * 1e3
        print("n in [%d,%d): %d calls; median ms add %.3f fit %.3f candidate %.3f deriv %.3f predict %.3f | max predict %.2f candidate %.2f fit %.2f"
              % (lo, hi, sel.sum(), md[1], md[2], md[3], md[4], md[5], mx[5], mx[3], mx[2]))
print("kernels:", {k: sum(1 for r in rows if r[7] == k) for k in set(r[7] for r in rows)}, "fits:", {k: sum(1 for r in rows if r[6] == k) for k in set(r[6] for r in rows)})
