"""Wall-clock of single C-ABI calls at strategy scale (what a proposal pays per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
for d, n in ((20, 3000), (30, 800), (10, 400)):
    rng = np.random.default_rng(0)
    lb, ub = -5.12 * np.ones(d), 5.12 * np.ones(d)
    X = lb + (ub - lb) * rng.random((n, d))
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
    s.add_points(X, np.sin(X.sum(1)))
    s._fit()
    h = s.handle
    for m in (1, 8, 2000, 100 * d):
        q = lb + (ub - lb) * rng.random((m, d))
        out = {}
        for name, fn in (("predict", lambda: h.predict(q, lb, ub)), ("predict_deriv", lambda: h.predict_deriv(q, lb, ub)),
                         ("score", lambda: h.score(q, lb, ub, X, True)), ("score+select", lambda: (h.score(q, lb, ub, X, True), h.select([1.0], 1, 1e-3)))):
            fn(); fn()
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            out[name] = (time.perf_counter() - t0) / 20 * 1e3
        print("d=%d n=%d m=%d: " % (d, n, m) + ", ".join("%s %.3f ms" % kv for kv in out.items()), flush=True)
