"""Teacher-forced replay of the strategy traces written by tools/make_strategy_traces.py (shared by the CPU oracle
test and the GPU parity test)."""
import os
import time

import numpy as np
import pytest

from conftest import GOLDEN, Problem, relmax


class GpuBackend:
    def __init__(self):
        import pysot_b200 as pb

        self.pb = pb
        self.candidate_dycors, self.candidate_srbf = pb.candidate_dycors, pb.candidate_srbf

    def surrogate(self, d, lb, ub):
        return self.pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=self.pb.CubicKernel(), tail=self.pb.LinearTail(d))


class OracleBackend:
    def __init__(self):
        from oracle import rbf_oracle as orc

        self.orc = orc
        self.candidate_dycors, self.candidate_srbf = orc.candidate_dycors, orc.candidate_srbf

    def surrogate(self, d, lb, ub):
        return self.orc.OracleRBF(d, lb, ub, "cubic", "linear")


def replay(fname, backend, max_calls=None, timer=None):
    """timer: optional dict; timer["seconds"] accumulates the wall-clock of add_points + the candidate_* call (what the
    trace's ref_call_seconds holds for the reference: its lazy refit happens inside the call)."""
    path = os.path.join(GOLDEN, fname)
    if not os.path.exists(path):
        pytest.skip("%s not generated" % fname)
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}  # NpzFile re-reads the archive member on every access
    d = int(g["dim"])
    prob = Problem(g["lb"], g["ub"], g["int_var"])
    kind = str(g["kind"])
    states = {int(c): i for i, c in enumerate(g["state_call"])}
    s = backend.surrogate(d, prob.lb, prob.ub)
    ncalls = len(g["n"]) if max_calls is None else min(max_calls, len(g["n"]))
    seg, have = -1, 0
    pend_off = np.concatenate(([0], np.cumsum(g["pend_cnt"])))
    pt_off = np.concatenate(([0], np.cumsum(g["num_pts"])))
    mismatches, ties, err = [], 0, {}
    for i in range(ncalls):
        if int(g["seg"][i]) != seg:  # restart: the strategy reset the surrogate (surrogate_strategy.py:228)
            s.reset()
            seg, have = int(g["seg"][i]), 0
        n = int(g["n"][i])
        t_start = time.perf_counter()
        if n > have:
            s.add_points(g["allX"][seg + have:seg + n], g["allfX"][seg + have:seg + n])
            have = n
        X, fX = g["allX"][seg:seg + n], g["allfX"][seg:seg + n]
        Xpend = g["pend"][pend_off[i]:pend_off[i + 1]]
        p = int(g["num_pts"][i])
        w = list(g["weights"][pt_off[i]:pt_off[i + 1]])
        if i in states:
            k = states[i]
            np.random.set_state(("MT19937", g["state_key"][k], int(g["state_pos"][k]), int(g["state_gauss"][k]),
                                 float(g["state_cached"][k])))
        if kind == "dycors":
            xb = g["xbest"][i]
            pts = backend.candidate_dycors(num_pts=p, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=w,
                                      prob_perturb=float(g["prob_perturb"][i]), Xpend=Xpend,
                                      sampling_radius=float(g["sampling_radius"][i]), num_cand=int(g["num_cand"][i]),
                                      xbest=None if np.isnan(xb[0]) else xb.copy())
        else:
            pts = backend.candidate_srbf(num_pts=p, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=w, Xpend=Xpend,
                                    sampling_radius=float(g["sampling_radius"][i]), num_cand=int(g["num_cand"][i]))
        if timer is not None:
            timer["seconds"] = timer.get("seconds", 0.0) + time.perf_counter() - t_start
        ref = g["new_points"][pt_off[i]:pt_off[i + 1]]
        if not np.array_equal(pts, ref):
            first = int(np.argmax(np.any(pts != ref, axis=1)))
            if g["margins"][pt_off[i] + first] <= 1e-12:
                ties += 1
            else:
                mismatches.append((i, first, float(g["margins"][pt_off[i] + first])))
        if "deriv" in g:  # config #5 add-on: value and gradient at every proposed point
            sl = slice(pt_off[i], pt_off[i + 1])
            gd, gp = s.predict_deriv(ref), s.predict(ref)
            for key, got in (("deriv", gd), ("deriv_full", gd), ("pred", gp), ("pred_full", gp)):
                if not np.isnan(g[key][sl]).any():  # the fresh-fit columns are recorded for a subset of the calls
                    err[key] = max(err.get(key, 0.0), float(np.max(np.abs(got - g[key][sl]))))
    if err:  # max |delta| / max |reference| over the whole trace (the north-star parity metric)
        n_used = pt_off[ncalls]
        sd, sp = float(np.nanmax(np.abs(g["deriv_full"][:n_used]))), float(np.nanmax(np.abs(g["pred_full"][:n_used])))
        err = {"pred_vs_incremental": err["pred"] / sp, "deriv_vs_incremental": err["deriv"] / sd,
               "pred_vs_fresh_fit": err["pred_full"] / sp, "deriv_vs_fresh_fit": err["deriv_full"] / sd}
    return ncalls, mismatches, ties, err
