"""Where the e2e gap of the bench comes from: b200rbf_score on the config #4 shape with device-resident candidates,
pinned host candidates and pageable host candidates, for several chunk sizes (B200RBF_OPT_CHUNK_ROWS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysot_b200 as pb
from pysot_b200 import _lib

d, n, m = 50, 20000, 1 << 20
rng = np.random.default_rng(0)
lb, ub = np.zeros(d), np.ones(d)
X = rng.random((n, d)); fX = np.sin(X.sum(1))[:, None]
s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
s.add_points(X, fX); s._fit()
h = s.handle
cand = rng.random((m, d))
pin = torch.from_numpy(cand).pin_memory()
cand_pin = pin.numpy()
dev = pin.cuda()
def run(c, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h.score(c, lb, ub, X, True, m=m); h.select([0.3, 0.5, 0.8, 0.95], 4, 1e-3)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
print("device-resident: %.2f ms (kernel %.2f ms)" % (run(dev), h.last_score_ms))
for chunk in (0, 16384, 32768, 50000, 65536, 131072):
    if chunk: h.set_option(_lib.OPT_CHUNK_ROWS, chunk)
    print("chunk_rows %7s: pinned host %.2f ms (kernels %.2f ms)   pageable host %.2f ms" % (chunk or "default", run(cand_pin), h.last_score_ms, run(cand)))
