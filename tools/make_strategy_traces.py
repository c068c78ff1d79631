#!/usr/bin/env python3
"""Record proposal-by-proposal traces of the UNMODIFIED reference strategies (build container only).

    python tools/make_strategy_traces.py        ->  tests/golden/trace_{dycors,srbf,sop}.npz

The reference strategies (pySOT/strategy/*.py) run with the reference surrogate under the POAP stand-in
(tools/poap_standin.py) on BASELINE.json's strategy-level configs.  Every call the strategy makes to
candidate_dycors / candidate_srbf is intercepted and its arguments, the RNG state it started from (only when
something other than the previous call advanced it) and the points it returned are stored.  The GPU test
(tests/test_gpu_strategy_traces.py) replays the calls through pysot_b200 with the same arguments and RNG stream
("teacher forcing": the recorded points are fed back, so one tie cannot derail the rest of the comparison).

For every pick the margin between the best and the second best merit of the reference is stored as well, so the
test can tell a legitimate tie (<= 1e-12, north_star) from a wrong answer.

Configs (BASELINE.json):
  #1 DYCORSStrategy, Ackley 10-D, cubic + linear tail, SLHD 2(d+1), 500 evals, SerialController
  #2 SRBFStrategy, Levy 30-D, 8 asynchronous workers (deterministic emulation), 100*dim candidates, 2000 evals
  #5 SOPStrategy, Rastrigin 20-D, 8 centers, 8 asynchronous workers (proposals are only requested while a worker is free,
     see tools/poap_standin.py), predict_deriv at every proposed point; 3000 evals (with the
     numpy nd_sorting of pysot_b200/pareto.py bound in place of the reference's pure-Python one -- identical output --
     because the latter alone makes 3000 evals a multi-hour run)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import poap_standin  # noqa: E402

poap_standin.install()
from tools import ref_import  # noqa: E402

ref_import.load()
import pySOT.strategy.dycors_strategy as m_dycors  # noqa: E402
import pySOT.strategy.sop_strategy as m_sop  # noqa: E402
import pySOT.strategy.srbf_strategy as m_srbf  # noqa: E402
from pySOT.auxiliary_problems import candidate_dycors as ref_dycors  # noqa: E402
from pySOT.auxiliary_problems import candidate_srbf as ref_srbf  # noqa: E402
from pySOT.experimental_design import SymmetricLatinHypercube  # noqa: E402
from pySOT.optimization_problems import Ackley, Levy, Rastrigin  # noqa: E402
from pySOT.strategy import DYCORSStrategy, SOPStrategy, SRBFStrategy  # noqa: E402
from pySOT.surrogate import CubicKernel, LinearTail, RBFInterpolant  # noqa: E402

from oracle import rbf_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def same_state(a, b):
    return a is not None and b is not None and a[2] == b[2] and np.array_equal(a[1], b[1]) and a[3] == b[3] and a[4] == b[4]


class Recorder:
    def __init__(self, kind, ref_fn, want_deriv=False):
        self.kind, self.ref_fn, self.want_deriv = kind, ref_fn, want_deriv
        self.calls, self.states, self.last_post = [], [], None
        self.ref_seconds = []  # wall time of each unmodified reference call (fit + generation + merit)
        self.strategy = None
        # every row that ever was in the strategy's current-run arrays (_X, _fX), in order; a call refers to the
        # contiguous slice [start, start + n).  (_X is NOT a slice of X: evaluations proposed before a radius
        # change are added to X but not to _X, surrogate_strategy.py:426-428.)
        self.runX, self.runfX, self.run_start, self.run_len = [], [], 0, 0

    def __call__(self, **kw):
        pre = np.random.get_state()
        t0 = time.perf_counter()
        pts = self.ref_fn(**kw)  # includes the lazy (incremental) refit inside surrogate.predict
        self.ref_seconds.append(time.perf_counter() - t0)
        post = np.random.get_state()
        # oracle replay from the same state: candidate set, picks and merit margins
        np.random.set_state(pre)
        prob, sur = kw["opt_prob"], kw["surrogate"]
        if self.kind == "dycors":
            cand = orc.gen_dycors(prob, kw["X"], kw["fX"], kw["prob_perturb"], kw.get("sampling_radius", 0.2),
                                  kw.get("subset"), kw.get("num_cand"), kw.get("xbest"))
        else:
            cand = orc.gen_srbf(prob, kw["X"], kw["fX"], kw.get("sampling_radius", 0.2), kw.get("subset"),
                                kw.get("num_cand"))
        assert same_state(np.random.get_state(), post), "oracle generator consumed the RNG differently"
        tr = {}
        opts = orc.weighted_distance_merit(kw["num_pts"], sur, kw["X"], kw["fX"], cand, kw["weights"], kw.get("Xpend"),
                                           kw.get("dtol", 1e-3), trace=tr)
        assert np.array_equal(opts, pts), "oracle and reference disagree"
        margins = []
        for mer in tr["merit"]:
            v = np.sort(mer.ravel())
            margins.append(v[1] - v[0] if np.isfinite(v[1]) else np.inf)
        np.random.set_state(post)
        if not same_state(pre, self.last_post):
            self.states.append((len(self.calls), pre))
        self.last_post = post
        Xc, fXc = kw["X"], kw["fX"]
        n = Xc.shape[0]
        cur = np.array(self.runX[self.run_start:self.run_start + min(n, self.run_len)]).reshape(-1, prob.dim)
        if n < self.run_len or not np.array_equal(cur, Xc[:self.run_len]):  # restart: a new run begins
            self.run_start, self.run_len = len(self.runX), 0
        for r in range(self.run_len, n):
            self.runX.append(Xc[r].copy())
            self.runfX.append(fXc[r].copy())
        self.run_len = n
        rec = dict(n=n, seg=self.run_start, Xpend=np.array(kw.get("Xpend")),
                   weights=np.array(kw["weights"], dtype=float), num_pts=kw["num_pts"],
                   prob_perturb=kw.get("prob_perturb", np.nan), sampling_radius=kw.get("sampling_radius", 0.2),
                   num_cand=kw.get("num_cand") or 100 * prob.dim,
                   xbest=np.array(kw["xbest"]) if kw.get("xbest") is not None else np.full(prob.dim, np.nan),
                   new_points=pts.copy(), margins=np.array(margins), picked=tr["picked"].copy())
        if self.want_deriv:
            rec["deriv"] = sur.predict_deriv(pts)  # the strategy's surrogate: incrementally extended LU (rbf.py:132-161)
            rec["pred"] = sur.predict(pts)
            if n <= 600 or len(self.calls) % 8 == 0:  # a fresh O(n^3) reference fit per call: every 8th call once n > 600
                fresh = RBFInterpolant(dim=prob.dim, lb=prob.lb, ub=prob.ub, kernel=CubicKernel(), tail=LinearTail(prob.dim))
                fresh.add_points(sur.X, sur.fX)     # the same points through the reference's initial-fit branch
                rec["deriv_full"] = fresh.predict_deriv(pts)
                rec["pred_full"] = fresh.predict(pts)
            else:
                rec["deriv_full"] = np.full_like(rec["deriv"], np.nan)
                rec["pred_full"] = np.full_like(rec["pred"], np.nan)
        self.calls.append(rec)
        return pts

    def save(self, fname, prob, extra):
        c = self.calls
        d = prob.dim
        st = self.strategy
        pend = [r["Xpend"].reshape(-1, d) for r in c]
        out = dict(
            kind=np.array(self.kind), dim=np.array(d), lb=prob.lb, ub=prob.ub, int_var=np.array(prob.int_var, dtype=np.int64),
            allX=np.array(self.runX), allfX=np.array(self.runfX).reshape(-1, 1), n=np.array([r["n"] for r in c]), seg=np.array([r["seg"] for r in c]),
            pend_cnt=np.array([p.shape[0] for p in pend]), pend=np.vstack(pend) if pend else np.empty((0, d)),
            num_pts=np.array([r["num_pts"] for r in c]), weights=np.concatenate([r["weights"] for r in c]),
            prob_perturb=np.array([r["prob_perturb"] for r in c]), sampling_radius=np.array([r["sampling_radius"] for r in c]),
            num_cand=np.array([r["num_cand"] for r in c]), xbest=np.vstack([r["xbest"] for r in c]),
            new_points=np.vstack([r["new_points"] for r in c]), margins=np.concatenate([r["margins"] for r in c]),
            picked=np.concatenate([r["picked"] for r in c]),
            state_call=np.array([s[0] for s in self.states]), state_key=np.vstack([s[1][1] for s in self.states]),
            state_pos=np.array([s[1][2] for s in self.states]), state_gauss=np.array([s[1][3] for s in self.states]),
            state_cached=np.array([s[1][4] for s in self.states]), ref_call_seconds=np.array(self.ref_seconds), **extra)
        if self.want_deriv:
            for key in ("deriv", "pred", "deriv_full", "pred_full"):
                out[key] = np.vstack([r[key] for r in c])
            cf = [a for a in c if not np.isnan(a["pred_full"]).any()]
            dn = max(np.max(np.abs(a["deriv"] - a["deriv_full"])) / np.max(np.abs(a["deriv_full"])) for a in cf)
            pn = max(np.max(np.abs(a["pred"] - a["pred_full"])) / np.max(np.abs(a["pred_full"])) for a in cf)
            print("reference incremental-vs-full noise: predict %.2e, predict_deriv %.2e (worst call)" % (pn, dn))
            out["ref_noise_pred"], out["ref_noise_deriv"] = np.array(pn), np.array(dn)
        np.savez_compressed(os.path.join(OUT, fname), **out)
        size = os.path.getsize(os.path.join(OUT, fname)) / 1e6
        print("%s: %d calls, %d stored RNG states, %d evals, best %.4f, min margin %.2e, %.2f MB"
              % (fname, len(c), len(self.states), len(st.X), float(st.fX.min()), float(out["margins"].min()), size))


def run(name, strategy_cls, module, attr, ref_fn, prob, max_evals, controller, seed=0, want_deriv=False, **skw):
    rec = Recorder("dycors" if attr == "candidate_dycors" else "srbf", ref_fn, want_deriv)
    setattr(module, attr, rec)
    try:
        np.random.seed(seed)
        d = prob.dim
        rbf = RBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub, kernel=CubicKernel(), tail=LinearTail(d))
        slhd = SymmetricLatinHypercube(dim=d, num_pts=2 * (d + 1))
        strat = strategy_cls(max_evals=max_evals, opt_prob=prob, exp_design=slhd, surrogate=rbf, asynchronous=True, **skw)
        rec.strategy = strat
        controller.strategy = strat
        t0 = time.perf_counter()
        controller.run()
        dt = time.perf_counter() - t0
        # the complete evaluation history of the run (strategy.X / fX, completion order): what a live run of the same
        # strategy with the GPU surrogate must reproduce (tests/test_gpu_live_strategies.py)
        rec.save("trace_%s.npz" % name, prob, dict(reference_seconds=np.array(dt), max_evals=np.array(max_evals),
                                                   hist_X=np.array(strat.X), hist_fX=np.array(strat.fX)))
    finally:
        setattr(module, attr, ref_fn)


if __name__ == "__main__":
    which = sys.argv[1:] or ["dycors", "srbf", "sop"]
    if "dycors" in which:
        p = Ackley(dim=10)
        run("dycors", DYCORSStrategy, m_dycors, "candidate_dycors", ref_dycors, p, 500, poap_standin.SerialController(p.eval))
    if "srbf" in which:
        p = Levy(dim=30)
        run("srbf", SRBFStrategy, m_srbf, "candidate_srbf", ref_srbf, p, 2000, poap_standin.SimAsyncController(p.eval, 8))
    if "sop" in which:
        p = Rastrigin(dim=20)
        # BASELINE config #5 at its stated length.  The reference's pure-Python nd_sorting (utils.py:379-484) alone
        # would take hours over 3000 completions (115 s over the first 568); pysot_b200.pareto.nd_sorting returns the
        # identical ranking (tests/test_host_logic.py, every prefix of this history) and is bound in its place here.
        # Everything else -- strategy, surrogate, candidate search -- is the unmodified reference.
        from pysot_b200.pareto import nd_sorting as fast_nd_sorting
        m_sop.nd_sorting = fast_nd_sorting
        run("sop", SOPStrategy, m_sop, "candidate_dycors", ref_dycors, p, 3000, poap_standin.SimAsyncController(p.eval, 8),
            want_deriv=True, ncenters=8)
