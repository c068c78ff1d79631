"""RBF fits at BASELINE config #3's sizes (TPS, d=10) -- timing and the target of the ncu launch list.
usage: fit_profile.py N reps [lookahead 0/1] [gemm_stages 0/2/3] [method 0 auto / 1 LU / 2 Cholesky]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb
from pysot_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
la = int(sys.argv[3]) if len(sys.argv) > 3 else 1
stages = int(sys.argv[4]) if len(sys.argv) > 4 else 0
method = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rng = np.random.default_rng(0)
X = rng.random((n, 10))
fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
s = pb.GpuRBFInterpolant(dim=10, lb=np.zeros(10), ub=np.ones(10), kernel=pb.TPSKernel(), tail=pb.LinearTail(10))
s.add_points(X, fX)
s.handle.set_option(_lib.OPT_LU_LOOKAHEAD, la)
s.handle.set_option(_lib.OPT_GEMM_STAGES, stages)
s.handle.set_option(_lib.OPT_FIT_METHOD, method)
cs = []
for r in range(reps):
    s.updated = False
    t0 = time.perf_counter()
    s._fit()
    cs.append(s.c.copy())
    print("fit n=%d lookahead=%d stages=%d method=%s: device %.4f s, host-to-host %.4f s, launches %d, |c|max %.3e"
          % (n, la, stages, s.handle.last_fit_method, s.fit_seconds, time.perf_counter() - t0, s.handle.launches,
             np.abs(s.c).max()))
print("repeatability max|dc| = %.3e ; interp residual %.3e" % (max(np.abs(c - cs[0]).max() for c in cs),
      np.max(np.abs(s.predict(X[:1000]) + s.eta * s.c[11:1011] - fX[:1000]))))
