"""INT8 tensor-core score kernel (score_i8.cu): parity against the DMMA kernel and the oracle, then throughput at the
bench shape.  usage: i8_check.py [quick]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysot_b200 as pb
from pysot_b200 import _lib
from oracle import rbf_oracle as orc

def relmax(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

for kernel, d, n, m in (("cubic", 50, 1500, 20000), ("tps", 10, 700, 9000), ("cubic", 33, 300, 5000), ("linear", 7, 130, 3000)):
    rng = np.random.default_rng(d)
    lb, ub = -np.ones(d) * 2, np.ones(d) * 3
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(((X - lb) / (ub - lb)).sum(1) * 2)[:, None]
    cand = lb + (ub - lb) * rng.random((m, d))
    cand[:50] = X[:50]                      # coincident with centers
    cand[50:100] = np.clip(X[50:100] + 1e-6 * (ub - lb) * rng.standard_normal((50, d)), lb, ub)
    Xpend = lb + (ub - lb) * rng.random((3, d))
    K = {"cubic": pb.CubicKernel, "tps": pb.TPSKernel, "linear": pb.LinearKernel}[kernel]
    T = pb.ConstantTail if kernel == "linear" else pb.LinearTail
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=K(), tail=T(d))
    s.add_points(X, fX)
    w = [0.3, 0.5, 0.8, 0.95]
    i0, m0, f0, d0, _ = pb.merit_select_indices(4, s, X, cand, w, Xpend, want_scores=True)
    s.handle.set_option(_lib.OPT_I8_DIST, 1)
    i1, m1, f1, d1, _ = pb.merit_select_indices(4, s, X, cand, w, Xpend, want_scores=True)
    o = orc.OracleRBF(d, lb, ub, kernel, "constant" if kernel == "linear" else "linear")
    o.add_points(X, fX)
    tr = {}
    orc.weighted_distance_merit(4, o, X, fX, cand.copy(), w, Xpend, 1e-3, trace=tr)
    fo, do = tr["fvals_raw"].ravel(), tr["dmerit0"].ravel()
    print("%s d=%d n=%d m=%d: int8 vs oracle fvals %.2e dmerit %.2e picks %s | dmma vs oracle fvals %.2e dmerit %.2e | oracle picks %s int8 %s"
          % (kernel, d, n, m, relmax(f1, fo), relmax(d1, do), "same" if i1.tolist() == tr["picked"].tolist() else "DIFF",
             relmax(f0, fo), relmax(d0, do), tr["picked"].tolist(), i1.tolist()), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(0)
d, n, m = 50, 20000, 1 << 20
rng = np.random.default_rng(0)
X = rng.random((n, d))
s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d))
s.add_points(X, np.sin(3 * X.sum(1)))
s.set_coefficients(rng.standard_normal(n + d + 1))
cand = torch.rand((m, d), dtype=torch.float64, device="cuda")
for mode in (0, 1):
    s.handle.set_option(_lib.OPT_I8_DIST, mode)
    for _ in range(2):
        idx, _ = pb.merit_select_indices(1, s, X, cand, [0.5], None)
    ms = []
    for _ in range(3):
        idx, _ = pb.merit_select_indices(1, s, X, cand, [0.5], None)
        ms.append(s.handle.last_score_ms)
    print("bench shape d=50 n=20000 m=2^20, %s kernel: %.2f ms per launch = %.3e pairs/s, pick %s" % ("int8 " if mode else "DMMA", min(ms), m * n / min(ms) * 1e3, idx.tolist()))
