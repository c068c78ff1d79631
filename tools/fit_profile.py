"""One RBF fit at BASELINE config #3's largest size (TPS, d=10, N=20000) -- target of the ncu launch list."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200 as pb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
X = rng.random((n, 10))
fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
s = pb.GpuRBFInterpolant(dim=10, lb=np.zeros(10), ub=np.ones(10), kernel=pb.TPSKernel(), tail=pb.LinearTail(10))
s.add_points(X, fX)
for r in range(reps):
    s.updated = False
    t0 = time.perf_counter()
    s._fit()
    print("fit n=%d: device %.4f s, host-to-host %.4f s, launches so far %d" % (n, s.fit_seconds, time.perf_counter() - t0, s.handle.launches))
print(s.handle.fp64_peaks(100.0))
