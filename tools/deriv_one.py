"""One predict_deriv configuration on device tensors (target of the ncu captures of the gradient kernels).
usage: deriv_one.py kernel d n log2m [mma 0/1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysot_b200 as pb
from pysot_b200 import _lib

kernel, d, n, lg = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mma = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(0)
X = rng.random((n, d))
s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d), kernel=pb.TPSKernel() if kernel == "tps" else pb.CubicKernel(), tail=pb.LinearTail(d))
s.add_points(X, np.sin(3 * X.sum(1)))
s.set_coefficients(rng.standard_normal(n + d + 1))
s.handle.set_option(_lib.OPT_MMA_DIST, mma)
m = 1 << lg
q = torch.rand((m, d), dtype=torch.float64, device="cuda")
out = torch.empty((m, d), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
s.handle.set_stream(st.cuda_stream)
for _ in range(2):
    s.handle.predict_deriv(q, s.lb, s.ub, out=out, m=m)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
s.handle.predict_deriv(q, s.lb, s.ub, out=out, m=m)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("predict_deriv %s d=%d n=%d m=2^%d mma=%d: %.2f ms, %.3e pairs/s" % (kernel, d, n, lg, mma, ms, m * n / ms * 1e3))
