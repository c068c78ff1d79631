"""predict_deriv throughput (device tensors, CUDA events): the DMMA gradient kernel against the direct tile kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysot_b200 as pb
from pysot_b200 import _lib

for kernel, d, n in (("tps", 10, 20000), ("cubic", 10, 20000), ("cubic", 20, 3000), ("cubic", 20, 20000), ("cubic", 50, 20000), ("tps", 50, 20000)):
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d), kernel=pb.TPSKernel() if kernel == "tps" else pb.CubicKernel(), tail=pb.LinearTail(d))
    s.add_points(X, np.sin(3 * X.sum(1)))
    s.set_coefficients(rng.standard_normal(n + d + 1))
    m = 1 << 20 if n >= 20000 else 1 << 22
    q = torch.rand((m, d), dtype=torch.float64, device="cuda")
    out = torch.empty((m, d), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    s.handle.set_stream(st.cuda_stream)
    res = {}
    for mode in (1, 0):
        s.handle.set_option(_lib.OPT_MMA_DIST, mode)
        s.handle.predict_deriv(q, s.lb, s.ub, out=out, m=m)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            s.handle.predict_deriv(q, s.lb, s.ub, out=out, m=m)
        e1.record(st)
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 3
    s.handle.set_stream(None)
    print("%s d=%d n=%d m=%d: tensor path %.2f ms = %.3e pairs/s (%.1f TFLOP/s algorithmic 5d+6) | direct tile kernel %.2f ms = %.3e pairs/s"
          % (kernel, d, n, m, res[1], m * n / res[1] * 1e3, m * n / res[1] * 1e3 * (5 * d + 6) / 1e12, res[0], m * n / res[0] * 1e3))
