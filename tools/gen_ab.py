"""Draw candidate sets on the device for a grid of (kind, d, subset, integer variables, m) and save them: run once per
library build (B200RBF_LIB) and compare the files -- the generator is counter-based, so a change of the launch geometry
must not change a single bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pysot_b200 as pb
from conftest import Problem
out = {}
rng = np.random.default_rng(0)
for d in (1, 5, 30, 50, 70):
    lb, ub = -1.0 - rng.random(d), 2.0 + 3.0 * rng.random(d)
    ints = tuple(range(0, d, 3)) if d >= 5 else ()
    for iv in ((), ints):
        prob = Problem(np.floor(lb) if iv else lb, np.ceil(ub) if iv else ub, iv)
        X = prob.lb + (prob.ub - prob.lb) * rng.random((2 * d + 4, d))
        if iv:
            X[:, list(iv)] = np.round(X[:, list(iv)])
        f = np.sin(X.sum(1))
        s = pb.GpuRBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub)
        s.add_points(X, f)
        for kind in ("dycors", "srbf", "uniform"):
            for subset in (None, list(range(0, d, 2))):
                for m in (1, 17, 3000, 100001):
                    for pp in (1.0, 0.05):
                        c = pb.generate_on_device(kind, prob, s, X, f, prob_perturb=pp, sampling_radius=0.2, subset=subset, num_cand=m, seed=12345 + m)
                        out["%d_%d_%s_%s_%d_%g" % (d, len(iv), kind, "all" if subset is None else "half", m, pp)] = c.cpu().numpy()
np.savez(sys.argv[1], **out)
print(len(out), "sets")
