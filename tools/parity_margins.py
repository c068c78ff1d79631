"""Headroom of the GPU parity tests: worst relative errors (max|delta| / max|ref|) over the committed golden fit /
predict / predict_deriv cases, against the 1e-10 bar of the north star."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pysot_b200 as pb
from conftest import Cases, relmax

K = {"cubic": "CubicKernel", "tps": "TPSKernel", "linear": "LinearKernel"}
T = {"linear": "LinearTail", "constant": "ConstantTail"}
cases = Cases("fit_cases.npz")
rows = []
for name in cases.names:
    g = cases.get(name)
    kernel, tail = name.split("_")[:2]
    d = g["X"].shape[1]
    s = pb.GpuRBFInterpolant(dim=d, lb=g["lb"], ub=g["ub"], kernel=getattr(pb, K[kernel])(), tail=getattr(pb, T[tail])(d))
    s.add_points(g["X"], g["fX"])
    rows.append((relmax(s.predict(g["q"]), g["pred"]), relmax(s.predict_deriv(g["dq"]), g["deriv"]), name, g["X"].shape[0],
                 s.handle.last_fit_method))
rows.sort(reverse=True)
print("%d cases; worst predict %.2e, worst predict_deriv %.2e (bar 1e-10)" % (len(rows), max(r[0] for r in rows), max(r[1] for r in rows)))
for r in rows[:6]:
    print("  %-28s n=%4d %-8s predict %.2e  deriv %.2e" % (r[2], r[3], r[4], r[0], r[1]))
