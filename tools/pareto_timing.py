"""Host time of SOP's non-dominated sorting over the evaluation history of the config #5 trace: the reference's
pure-Python nd_sorting against pysot_b200.pareto.nd_sorting (same output).  One call per completed evaluation, as in
sop_strategy.py:385-403 (objectives: f(x) and minus the distance to the nearest other point, nmax = 100)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.spatial.distance import cdist
from pysot_b200.pareto import nd_sorting as fast

ref = None
if os.path.isdir("/root/reference/pySOT"):
    from tools import ref_import
    ref_import.load()
    from pySOT.utils import nd_sorting as ref

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trace_sop.npz"))
X, fX, lb, ub = g["allX"], g["allfX"].ravel(), g["lb"], g["ub"]
N = min(len(X), 600)
U = (X[:N] - lb) / (ub - lb)
D = cdist(U, U)
t_ref = t_fast = 0.0
same = True
for n in range(25, N + 1):
    d = D[:n, :n].copy()
    d[d == 0.0] = np.inf
    F = np.vstack((fX[:n], -d.min(axis=1)))
    t0 = time.perf_counter(); b = fast(F, 100); t_fast += time.perf_counter() - t0
    if ref is not None:
        t0 = time.perf_counter(); a = ref(F, 100); t_ref += time.perf_counter() - t0
        same &= bool(np.array_equal(a, b))
print("nd_sorting over %d completions (n = 25 .. %d): reference %.2f s, numpy %.3f s, identical ranks: %s"
      % (N - 24, N, t_ref, t_fast, same if ref is not None else "reference not present"))
