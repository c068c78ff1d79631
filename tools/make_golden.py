#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tools/make_golden.py

For every case the reference (pySOT.surrogate.RBFInterpolant,
pySOT.auxiliary_problems.*) and the oracle (oracle/rbf_oracle.py) are run on the
same seeded inputs; the script asserts that they agree (bit for bit wherever
the same compiled routines are called in the same order) and stores the
REFERENCE's outputs.  Inputs are stored too, so the fixtures are self-contained.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ref_import  # noqa: E402

pySOT = ref_import.load()
from pySOT.auxiliary_problems import candidate_dycors, candidate_srbf, candidate_uniform  # noqa: E402
from pySOT.auxiliary_problems.candidate_srbf import weighted_distance_merit  # noqa: E402
from pySOT.optimization_problems import Ackley  # noqa: E402
from pySOT.surrogate import (  # noqa: E402
    ConstantTail, CubicKernel, LinearKernel, LinearTail, RBFInterpolant, TPSKernel, median_capping,
)

from oracle import rbf_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

KERNELS = {"cubic": CubicKernel, "tps": TPSKernel, "linear": LinearKernel}
TAILS = {"linear": LinearTail, "constant": ConstantTail}
COMBOS = [("cubic", "linear"), ("tps", "linear"), ("linear", "linear"), ("linear", "constant")]


def ref_rbf(dim, lb, ub, kernel, tail, eta=1e-6, output_transformation=None):
    return RBFInterpolant(dim=dim, lb=lb, ub=ub, kernel=KERNELS[kernel](), tail=TAILS[tail](dim), eta=eta,
                          output_transformation=output_transformation)


def make_box(rng, d, kind):
    if kind == "unit":
        return np.zeros(d), np.ones(d)
    if kind == "iso":  # Ackley-like isotropic box
        return -15.0 * np.ones(d), 20.0 * np.ones(d)
    lb = rng.uniform(-10, 0, d)
    return lb, lb + rng.uniform(0.5, 30, d)


def objective(X, lb, ub):
    U = (X - lb) / (ub - lb)
    return (np.sin(3.0 * U.sum(axis=1)) + U[:, 0] ** 2 + 0.1 * np.cos(5.0 * U).sum(axis=1))[:, None]


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ------------------------------------------------------------------ A. fit / predict / predict_deriv
def gen_fit_cases():
    rng = np.random.default_rng(20261019)
    store, names = {}, []
    worst = 0.0
    for kernel, tail in COMBOS:
        for d in (1, 2, 10, 30, 50):
            t = d + 1 if tail == "linear" else 1
            for n, box in ((t, "unit"), (max(t + 9, 60), "aniso"), (400, "iso")):
                if n == t and tail == "constant":
                    n = 3  # n == 1 gives a singular all-equal system; keep a tiny but solvable one
                lb, ub = make_box(rng, d, box)
                X = lb + (ub - lb) * rng.random((n, d))
                fX = objective(X, lb, ub)
                q = lb + (ub - lb) * rng.random((48, d))
                q[:4] = X[:4] if n >= 4 else q[:4]  # include evaluated points (r = 0 paths)
                dq = q[:12]
                r = ref_rbf(d, lb, ub, kernel, tail)
                r.add_points(X, fX)
                pred = r.predict(q)
                der = r.predict_deriv(dq)
                o = orc.OracleRBF(d, lb, ub, kernel, tail)
                o.add_points(X, fX)
                po, do = o.predict(q), o.predict_deriv(dq)
                assert np.array_equal(o.c, r.c), (kernel, tail, d, n)
                assert np.array_equal(po, pred) and np.array_equal(do, der), (kernel, tail, d, n)
                name = f"{kernel}_{tail}_d{d}_n{n}"
                names.append(name)
                for k, v in dict(X=X, fX=fX, lb=lb, ub=ub, q=q, dq=dq, c=r.c, pred=pred, deriv=der).items():
                    store[f"{name}/{k}"] = v
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "fit_cases.npz"), **store)
    print("fit_cases:", len(names), "cases; oracle == reference bit for bit")


# ------------------------------------------------------------------ B. reference test_rbf scenario (incremental refits)
def gen_incremental():
    n = 30
    xv, yv = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n))
    X = np.hstack((xv.reshape(-1, 1), yv.reshape(-1, 1)))
    fX = (X[:, 1] * np.sin(X[:, 0]) + X[:, 0] * np.cos(X[:, 1]))[:, None]
    np.random.seed(0)
    Xs = np.random.rand(10, 2)
    r = RBFInterpolant(dim=2, lb=np.zeros(2), ub=np.ones(2), eta=1e-6)
    o = orc.OracleRBF(2, np.zeros(2), np.ones(2))
    r.add_points(X, fX)
    o.add_points(X, fX)
    full_pred, full_der = r.predict(Xs), r.predict_deriv(Xs)
    assert np.array_equal(o.predict(Xs), full_pred)
    r.reset()
    o.reset()
    steps = []
    for i in range(9):
        sl = slice(i * 100, (i + 1) * 100)
        r.add_points(X[sl], fX[sl])
        o.add_points(X[sl], fX[sl])
        steps.append(r.predict(Xs))
        assert np.array_equal(o.predict(Xs), steps[-1]), i
    inc_der = r.predict_deriv(Xs)
    assert np.array_equal(o.predict_deriv(Xs), inc_der)
    np.savez_compressed(os.path.join(OUT, "incremental.npz"), X=X, fX=fX, Xs=Xs, full_pred=full_pred,
                        full_deriv=full_der, step_pred=np.array(steps), inc_deriv=inc_der, c_inc=r.c)
    print("incremental: incremental-vs-full predict diff (reference's own noise floor): %.2e" %
          relmax(steps[-1], full_pred))


# ------------------------------------------------------------------ C. median capping (reference test_capped)
def gen_capped():
    rng = np.random.default_rng(7)
    d, n = 3, 120
    lb, ub = -2 * np.ones(d), 3 * np.ones(d)
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.exp(3 * rng.random((n, 1)))
    q = lb + (ub - lb) * rng.random((20, d))
    r = ref_rbf(d, lb, ub, "cubic", "linear", output_transformation=median_capping)
    r.add_points(X, fX)
    o = orc.OracleRBF(d, lb, ub, output_transformation=orc.median_capping)
    o.add_points(X, fX)
    pred, der = r.predict(q), r.predict_deriv(q)
    assert np.array_equal(o.predict(q), pred) and np.array_equal(o.predict_deriv(q), der)
    np.savez_compressed(os.path.join(OUT, "capped.npz"), X=X, fX=fX, lb=lb, ub=ub, q=q, pred=pred, deriv=der)
    print("capped: ok")


# ------------------------------------------------------------------ D. weighted_distance_merit
MERIT_CASES = [
    # name, kernel, tail, d, n, m, k_pend, box, weights, num_pts, dtol
    ("cubic_d10_p1", "cubic", "linear", 10, 120, 1000, 0, "iso", [0.3], 1, 1e-3),
    ("cubic_d10_p4_pend", "cubic", "linear", 10, 120, 1000, 3, "iso", [0.3, 0.5, 0.8, 0.95], 4, 1e-3),
    ("tps_d30_p4_aniso", "tps", "linear", 30, 200, 1500, 7, "aniso", [0.3, 0.5, 0.8, 0.95], 4, 1e-3),
    ("cubic_d50_p4_unit", "cubic", "linear", 50, 300, 2048, 7, "unit", [0.3, 0.5, 0.8, 0.95], 4, 1e-3),
    ("linear_const_d2_p2", "linear", "constant", 2, 40, 700, 2, "aniso", [0.5, 0.95], 2, 1e-3),
    ("linear_lin_d5_p3", "linear", "linear", 5, 64, 513, 0, "unit", [0.95, 0.3, 0.5], 3, 1e-3),
    ("sop_w1_d20", "cubic", "linear", 20, 150, 2000, 0, "iso", [1.0], 1, 1e-3),
    ("w0_d4_p3", "cubic", "linear", 4, 30, 300, 1, "unit", [0.0, 0.0, 0.3], 3, 1e-3),
    ("bigdtol_d3_p4", "cubic", "linear", 3, 50, 400, 0, "unit", [0.3, 0.5, 0.8, 0.95], 4, 0.15),
    ("allmasked_d2_p2", "cubic", "linear", 2, 30, 200, 0, "unit", [0.3, 0.5], 2, 10.0),
    ("d1_p2", "cubic", "linear", 1, 12, 257, 1, "iso", [0.25, 0.75], 2, 1e-3),
]


def gen_merit_cases():
    rng = np.random.default_rng(4242)
    store, names = {}, []
    for name, kernel, tail, d, n, m, kp, box, weights, p, dtol in MERIT_CASES:
        lb, ub = make_box(rng, d, box)
        X = lb + (ub - lb) * rng.random((n, d))
        fX = objective(X, lb, ub)
        Xpend = lb + (ub - lb) * rng.random((kp, d)) if kp else None
        xbest = X[np.argmin(fX)]
        # DYCORS-shaped candidates around xbest + some uniform ones + a few duplicates of evaluated points
        mask = rng.random((m, d)) < 0.4
        cand = np.where(mask, np.clip(xbest + 0.2 * (ub - lb) * rng.standard_normal((m, d)), lb, ub), xbest)
        cand[m // 2:] = lb + (ub - lb) * rng.random((m - m // 2, d))
        cand[5] = X[3]
        cand[17] = X[min(7, n - 1)] + 1e-5 * (ub - lb)
        r = ref_rbf(d, lb, ub, kernel, tail)
        r.add_points(X, fX)
        o = orc.OracleRBF(d, lb, ub, kernel, tail)
        o.add_points(X, fX)
        ref_pts = weighted_distance_merit(num_pts=p, surrogate=r, X=X, fX=fX, cand=cand.copy(), weights=weights,
                                          Xpend=Xpend, dtol=dtol)
        tr = {}
        orc_pts = orc.weighted_distance_merit(p, o, X, fX, cand.copy(), weights, Xpend, dtol, trace=tr)
        assert np.array_equal(ref_pts, orc_pts), name
        names.append(name)
        rec = dict(X=X, fX=fX, lb=lb, ub=ub, cand=cand, weights=np.array(weights), dtol=np.array(dtol),
                   c=r.c, new_points=ref_pts, picked=tr["picked"], picked_merit=tr["picked_merit"],
                   fvals_raw=tr["fvals_raw"], dmerit0=tr["dmerit0"], kernel=np.array(kernel), tail=np.array(tail))
        if Xpend is not None:
            rec["Xpend"] = Xpend
        for k, v in rec.items():
            store[f"{name}/{k}"] = v
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "merit_cases.npz"), **store)
    print("merit_cases:", len(names), "cases; oracle new_points == reference")


# ------------------------------------------------------------------ E. candidate generators (global numpy RNG order)
class _IntAckley(Ackley):
    def __init__(self, dim, int_var):
        super().__init__(dim)
        self.int_var = np.array(int_var)
        self.cont_var = np.array([i for i in range(dim) if i not in int_var])


def gen_candidate_cases():
    store, names = {}, []
    rng = np.random.default_rng(99)
    specs = [
        ("dycors_d10", "dycors", Ackley(dim=10), dict(prob_perturb=0.35, num_cand=500), 4),
        ("dycors_d10_sop", "dycors", Ackley(dim=10), dict(prob_perturb=1.0, num_cand=300, xbest="row5"), 1),
        ("dycors_d6_int", "dycors", _IntAckley(6, [1, 4]), dict(prob_perturb=0.2, num_cand=400, sampling_radius=0.1), 2),
        ("dycors_d1", "dycors", Ackley(dim=1), dict(prob_perturb=0.5, num_cand=200), 1),
        ("srbf_d10", "srbf", Ackley(dim=10), dict(num_cand=500, sampling_radius=0.2), 4),
        ("srbf_d6_int", "srbf", _IntAckley(6, [0, 5]), dict(num_cand=300, sampling_radius=0.05), 2),
        ("uniform_d10", "uniform", Ackley(dim=10), dict(num_cand=500), 4),
        ("uniform_d6_subset", "uniform", Ackley(dim=6), dict(num_cand=300, subset=[0, 2, 3]), 2),
    ]
    weights = [0.3, 0.5, 0.8, 0.95]
    for name, kind, prob, kw, p in specs:
        d = prob.dim
        n = 4 * (d + 1)
        X = prob.lb + (prob.ub - prob.lb) * rng.random((n, d))
        if len(prob.int_var) > 0:
            X[:, prob.int_var] = np.round(X[:, prob.int_var])
        fX = np.array([[prob.eval(x)] for x in X])
        Xpend = prob.lb + (prob.ub - prob.lb) * rng.random((2, d))
        kw = dict(kw)
        if kw.get("xbest") == "row5":
            kw["xbest"] = X[5].copy()
        r = RBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub)
        r.add_points(X, fX)
        o = orc.OracleRBF(d, prob.lb, prob.ub)
        o.add_points(X, fX)
        fn = dict(dycors=candidate_dycors, srbf=candidate_srbf, uniform=candidate_uniform)[kind]
        ofn = dict(dycors=orc.candidate_dycors, srbf=orc.candidate_srbf, uniform=orc.candidate_uniform)[kind]
        ogen = dict(dycors=orc.gen_dycors, srbf=orc.gen_srbf, uniform=orc.gen_uniform)[kind]
        np.random.seed(12345)
        ref_pts = fn(num_pts=p, opt_prob=prob, surrogate=r, X=X, fX=fX, weights=weights[:p], Xpend=Xpend, **kw)
        np.random.seed(12345)
        orc_pts = ofn(num_pts=p, opt_prob=prob, surrogate=o, X=X, fX=fX, weights=weights[:p], Xpend=Xpend, **kw)
        assert np.array_equal(ref_pts, orc_pts), name
        np.random.seed(12345)
        gkw = {k: v for k, v in kw.items()}
        cand = ogen(prob, X, fX, **gkw)
        names.append(name)
        rec = dict(X=X, fX=fX, Xpend=Xpend, lb=prob.lb, ub=prob.ub, int_var=np.array(prob.int_var, dtype=np.int64),
                   cand=cand, new_points=ref_pts, weights=np.array(weights[:p]), kind=np.array(kind), seed=np.array(12345))
        for k, v in kw.items():
            rec["kw_" + k] = np.array(v)
        for k, v in rec.items():
            store[f"{name}/{k}"] = v
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "candidate_cases.npz"), **store)
    print("candidate_cases:", len(names), "cases; oracle == reference under np.random.seed")


# ------------------------------------------------------------------ F. reference known answer (tests/test_auxiliary_problems.py:8-31)
def gen_known_answer():
    np.random.seed(0)
    ackley = Ackley(dim=1)
    X = np.expand_dims([-15.0, -10, 0, 1, 20], axis=1)
    fX = np.array([ackley.eval(x) for x in X])[:, None]
    r = RBFInterpolant(dim=1, lb=ackley.lb, ub=ackley.ub)
    r.add_points(X, fX)
    x_u = candidate_uniform(num_pts=1, X=X, Xpend=None, fX=fX, num_cand=10000, surrogate=r, opt_prob=ackley,
                            weights=[0.25])
    x_s = candidate_srbf(num_pts=1, X=X, Xpend=None, fX=fX, num_cand=10000, surrogate=r, opt_prob=ackley,
                         weights=[0.25], sampling_radius=0.5)
    assert abs(x_u[0, 0] - 10.5) < 1e-2 and abs(x_s[0, 0] - 10.5) < 1e-2
    np.savez_compressed(os.path.join(OUT, "known_answer.npz"), X=X, fX=fX, lb=ackley.lb, ub=ackley.ub,
                        x_uniform=x_u, x_srbf=x_s)
    print("known_answer: uniform %.4f srbf %.4f (reference test pins 10.50 +- 1e-2)" % (x_u[0, 0], x_s[0, 0]))


if __name__ == "__main__":
    gen_fit_cases()
    gen_incremental()
    gen_capped()
    gen_merit_cases()
    gen_candidate_cases()
    gen_known_answer()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures: %.2f MB in %s" % (total / 1e6, OUT))
