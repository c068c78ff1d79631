#!/usr/bin/env python3
"""Golden vectors at the BENCHMARKED shapes, from the live reference (build container only; ~5 min, ~25 GB RSS).

    python tools/make_large_goldens.py [fit5k] [fit20k] [config4]   ->  tests/golden/large_*.npz

  large_fit_tps10_{5000,20000}.npz   BASELINE config #3: TPS + linear tail, d = 10, X ~ U[0,1]^10 seed 0,
        fX = sin(3 sum X) + X0^2 (exactly what bench.py's fit sweep fits).  The UNMODIFIED reference
        RBFInterpolant (rbf.py:110-130: cdist, lu_factor, two solve_triangular) is fitted once; stored: 512
        predictions and 32 gradients at seed-1 points, the coefficient vector, the reference's wall-clock seconds.
  large_config4.npz                  BASELINE config #4 as bench.py builds it: d = 50, n = 20 000 centers, cubic +
        linear tail, coefficients from the reference's own fit; a 2^17-row DYCORS candidate slice drawn through the
        reference's candidate_dycors RNG path (np.random.seed(0), prob_perturb 0.4, sigma 0.2 -- regenerated on the GPU
        box by pysot_b200.generate_dycors, which is bit-identical; only its SHA-256 is stored); 7 pending points;
        fvals / dmerit of weighted_distance_merit (candidate_srbf.py:35-40) computed with the reference's
        surrogate.predict and cdist over 4096-row slices (the full 2^17 x 20 000 matrices do not fit in memory; row
        results do not depend on the slicing), the picks for weights [0.3, 0.5, 0.8, 0.95] and the margin between
        the best and second-best merit of every pick.
"""
import hashlib
import os
import sys
import time

import numpy as np
import scipy.spatial.distance as ssd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ref_import  # noqa: E402

ref_import.load()
from pySOT.auxiliary_problems import candidate_dycors as ref_candidate_dycors  # noqa: E402
from pySOT.surrogate import CubicKernel, LinearTail, RBFInterpolant, TPSKernel  # noqa: E402

from oracle import rbf_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fit_case(n):
    d = 10
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
    s = RBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d), kernel=TPSKernel(), tail=LinearTail(d), eta=1e-6)
    s.add_points(X, fX)
    t0 = time.perf_counter()
    s._fit()
    secs = time.perf_counter() - t0
    q = np.random.default_rng(1).random((512, d))
    pred = s.predict(q)
    grad = s.predict_deriv(q[:32])
    # residual of the reference's own solution, for the record
    resid = float(np.max(np.abs(s.predict(X[:512]) + s.eta * s.c[d + 1:d + 1 + 512] - fX[:512])))
    np.savez_compressed(os.path.join(OUT, "large_fit_tps10_%d.npz" % n), n=np.array(n), d=np.array(d), seed=np.array(0),
                        q=q, pred=pred, grad=grad, c=s.c.ravel(), ref_fit_seconds=np.array(secs), ref_residual=np.array(resid),
                        cores=np.array(os.cpu_count()))
    print("fit n=%d: reference _fit %.2f s on %d cores, residual %.2e, max|pred| %.3f" % (n, secs, os.cpu_count(), resid,
                                                                                        float(np.abs(pred).max())))


class _Prob:
    def __init__(self, d):
        self.dim, self.lb, self.ub = d, np.zeros(d), np.ones(d)
        self.int_var, self.cont_var = np.array([], dtype=int), np.arange(d)


def config4_case(m=1 << 17):
    d, n = 50, 20000
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
    Xpend = np.random.default_rng(1).random((7, d))
    s = RBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d), kernel=CubicKernel(), tail=LinearTail(d), eta=1e-6)
    s.add_points(X, fX)
    t0 = time.perf_counter()
    s._fit()
    secs = time.perf_counter() - t0
    print("config4: reference fit d=50 n=20000 %.1f s" % secs)
    # the candidate set: the reference's candidate_dycors with its merit function intercepted
    grabbed = {}

    def grab(num_pts, surrogate, X, fX, cand, weights, Xpend=None, dtol=1e-3):
        grabbed["cand"] = cand
        return cand[:num_pts].copy()

    m_cd = sys.modules["pySOT.auxiliary_problems.candidate_dycors"]  # the package attribute of that name is the function
    saved = m_cd.weighted_distance_merit
    m_cd.weighted_distance_merit = grab
    try:
        np.random.seed(0)
        ref_candidate_dycors(num_pts=4, opt_prob=_Prob(d), surrogate=s, X=X, fX=fX, weights=[0.3, 0.5, 0.8, 0.95],
                             prob_perturb=0.4, Xpend=Xpend, sampling_radius=0.2, num_cand=m)
    finally:
        m_cd.weighted_distance_merit = saved
    cand = grabbed["cand"]
    assert cand.shape == (m, d)
    np.random.seed(0)
    again = orc.gen_dycors(_Prob(d), X, fX, 0.4, 0.2, None, m, None)
    assert np.array_equal(again, cand), "oracle generator differs from the reference"
    sha = hashlib.sha256(np.ascontiguousarray(cand).tobytes()).hexdigest()
    # candidate_srbf.py:35-40 over row slices with the reference's own predict / cdist
    P = np.vstack((X, Xpend))
    fv, dm = np.empty((m, 1)), np.empty((m, 1))
    t0 = time.perf_counter()
    from concurrent.futures import ThreadPoolExecutor

    def work(lo):
        hi = min(lo + 4096, m)
        dm[lo:hi] = np.amin(ssd.cdist(cand[lo:hi], P), axis=1, keepdims=True)
        fv[lo:hi] = s.predict(cand[lo:hi])

    with ThreadPoolExecutor(os.cpu_count()) as ex:
        list(ex.map(work, range(0, m, 4096)))
    print("config4: reference scoring of %d x %d: %.1f s" % (m, n, time.perf_counter() - t0))
    weights = [0.3, 0.5, 0.8, 0.95]
    # the selection loop (candidate_srbf.py:40-55) with margins
    f = orc.unit_rescale(fv.copy())
    dcur = dm.copy()
    picked, margins, merits = [], [], []
    for w in weights:
        merit = w * f + (1.0 - w) * (1.0 - orc.unit_rescale(np.copy(dcur)))
        merit[dcur < 1e-3] = np.inf
        jj = int(np.argmin(merit))
        v = np.partition(merit.ravel(), 1)[:2]
        picked.append(jj)
        merits.append(float(merit[jj, 0]))
        margins.append(float(v[1] - v[0]))
        f[jj] = np.inf
        dcur = np.minimum(dcur, ssd.cdist(cand, np.atleast_2d(cand[jj])))
    assert np.array_equal(orc.select_from_scores(cand, fv, dm, weights, 4, 1e-3), np.array(picked))
    np.savez_compressed(os.path.join(OUT, "large_config4.npz"), m=np.array(m), n=np.array(n), d=np.array(d),
                        cand_sha256=np.array(sha), c=s.c.ravel(), fvals=fv.ravel(), dmerit=dm.ravel(), Xpend=Xpend,
                        weights=np.array(weights), picked=np.array(picked), merit=np.array(merits),
                        margins=np.array(margins), new_points=cand[picked], ref_fit_seconds=np.array(secs),
                        cores=np.array(os.cpu_count()))
    print("config4: picked %s, margins %s, sha %s" % (picked, margins, sha[:16]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["fit5k", "fit20k", "config4"]
    if "fit5k" in which:
        fit_case(5000)
    if "fit20k" in which:
        fit_case(20000)
    if "config4" in which:
        config4_case()
