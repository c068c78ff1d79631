"""Randomised check that pysot_b200's batched host candidate generators draw bit-identical candidate sets (and leave
numpy's global stream in the same state) as the oracle's per-coordinate scipy.stats.truncnorm.rvs loop, which is the
reference's (candidate_srbf.py:91-116, candidate_dycors.py:55-89).  Prints the host time per call of both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pysot_b200.candidates as c
from oracle import rbf_oracle as orc


class P:
    def __init__(s, lb, ub, iv):
        s.lb, s.ub, s.dim, s.int_var = lb, ub, len(lb), iv
        s.cont_var = np.setdiff1d(np.arange(len(lb)), iv)


def run(cases=300, seed=0):
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(cases):
        d = int(rng.integers(1, 40)); lb = -rng.random(d) * 10; ub = lb + rng.random(d) * 20 + 0.1
        iv = (np.array(sorted(rng.choice(d, size=int(rng.integers(0, min(d, 3) + 1)), replace=False)), dtype=int)
              if it % 3 == 0 else np.array([], dtype=int))
        p = P(lb, ub, iv); n = int(rng.integers(2, 30)); X = lb + (ub - lb) * rng.random((n, d)); fX = rng.random((n, 1))
        if it % 4 == 0:
            X[np.argmin(fX)] = np.where(rng.random(d) < 0.5, lb, ub)  # xbest on the boundary: a == 0 or b == 0
        sr = float(rng.choice([0.2, 0.05, 0.0001, 1.5])); pp = float(rng.choice([1.0, 0.3, 0.01]))
        nc = int(rng.integers(1, 400))
        subset = None if it % 5 else list(rng.choice(d, size=max(1, d // 2), replace=False))
        dsub = subset if subset is None else sorted(range(len(subset)))
        sd = int(rng.integers(1 << 30))
        for kind in ("srbf", "dycors"):
            res = []
            for mod in (c, orc):
                np.random.seed(sd)
                try:
                    if kind == "srbf":
                        out = (mod.generate_srbf if mod is c else mod.gen_srbf)(p, X, fX, sr, subset, nc)
                    else:
                        out = (mod.generate_dycors if mod is c else mod.gen_dycors)(p, X, fX, pp, sr, dsub, nc)
                except Exception as e:  # both must fail the same way
                    out = repr(e)
                res.append((out, np.random.get_state()[1].copy()))
            (a, sa), (b, sb) = res
            same = (a == b) if isinstance(a, str) or isinstance(b, str) else np.array_equal(a, b)
            if not (same and np.array_equal(sa, sb)):
                bad += 1
                print("MISMATCH", it, kind, d, nc, subset)
    return bad


if __name__ == "__main__":
    assert c._fast_truncnorm_ok(), "batched truncnorm path is switched off on this scipy"
    print("mismatches:", run())
    rng = np.random.default_rng(1)
    p = P(-15 * np.ones(30), 20 * np.ones(30), np.array([], dtype=int)); X = rng.random((50, 30)); fX = rng.random((50, 1))
    for f, name, args in ((c.generate_srbf, "batched srbf", (0.2, None, 3000)), (orc.gen_srbf, "scipy loop srbf", (0.2, None, 3000)),
                          (c.generate_dycors, "batched dycors", (0.3, 0.2, None, 3000)),
                          (orc.gen_dycors, "scipy loop dycors", (0.3, 0.2, None, 3000))):
        t0 = time.perf_counter()
        for _ in range(50):
            f(p, X, fX, *args)
        print("%-18s %.2f ms per call (d=30, 3000 candidates)" % (name, (time.perf_counter() - t0) / 50 * 1e3))
