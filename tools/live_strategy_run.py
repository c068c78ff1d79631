#!/usr/bin/env python3
"""Run pySOT's UNMODIFIED strategies live, with the GPU surrogate dropped in (or the stock CPU surrogate).

    python tools/live_strategy_run.py --config {1,2,5} [--surrogate gpu|cpu] [--controller serial|simasync|thread]
                                      [--evals N] [--checkpoint] [--generator host|device]

Must run in a fresh interpreter: pySOT (the reference, pip-installed unmodified into baseline/_ref by
`__graft_entry__.build()`, or /root/reference in the build container) has to be importable BEFORE pysot_b200 is, so that
GpuRBFInterpolant becomes a subclass of pySOT.surrogate.RBFInterpolant (pysot_b200/interface.py) and passes the
strategies' `isinstance(surrogate, Surrogate)` check (surrogate_strategy.py:154).  POAP is not available offline;
tools/poap_standin.py provides the controller protocol (serial, deterministic 8-worker emulation, real threads).

BASELINE.json configs:
  1  DYCORSStrategy, Ackley 10-D, cubic + linear tail, SLHD 2(d+1), 500 evals, SerialController
  2  SRBFStrategy, Levy 30-D, 8 asynchronous workers, 100*dim candidates per proposal, 2000 evals
  5  SOPStrategy, Rastrigin 20-D, 8 centers, 8 asynchronous workers, predict_deriv at every proposed point, 3000 evals

With --surrogate gpu the run uses pysot_b200.GpuRBFInterpolant and pysot_b200.install(); nothing else differs from the
run that produced tests/golden/trace_*.npz (same seed, same controller), so with the serial / simasync controllers the
evaluation history must equal the stored `hist_X` / `hist_fX` row for row.  Prints one JSON line.
"""
import argparse
import hashlib
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def import_reference():
    """pySOT from baseline/_ref (travels to the GPU box) or the build container's /root/reference; POAP stand-in."""
    from tools import poap_standin

    poap_standin.install()
    if "pyDOE2" not in sys.modules:
        sys.modules["pyDOE2"] = types.ModuleType("pyDOE2")
    import numpy as np

    if not hasattr(np, "int"):  # sop_strategy.py:457, utils.py:468 use the alias numpy removed
        np.int = int
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("PYSOT_REFERENCE", "/root/reference")):
        if os.path.isdir(os.path.join(cand, "pySOT")):
            sys.path.insert(0, cand)
            import pySOT

            return pySOT, cand, poap_standin
    raise SystemExit("pySOT not found: run `python -c 'import __graft_entry__ as g; g.build()'` in the build container "
                     "(it installs the reference into baseline/_ref)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[1, 2, 5])
    ap.add_argument("--surrogate", default="gpu", choices=["gpu", "cpu"])
    ap.add_argument("--controller", default=None, choices=["serial", "simasync", "thread"])
    ap.add_argument("--evals", type=int, default=None)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--checkpoint", action="store_true", help="dill-checkpoint every evaluation, abort mid-run, resume")
    ap.add_argument("--generator", default="host", choices=["host", "device"])
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    pySOT, where, poap = import_reference()
    import numpy as np
    from pySOT.controller import CheckpointController
    from pySOT.experimental_design import SymmetricLatinHypercube
    from pySOT.optimization_problems import Ackley, Levy, Rastrigin
    from pySOT.strategy import DYCORSStrategy, SOPStrategy, SRBFStrategy
    from pySOT.surrogate import CubicKernel, LinearTail, RBFInterpolant, Surrogate

    cfg = {1: (DYCORSStrategy, Ackley(dim=10), 500, "serial", {}, "trace_dycors.npz"),
           2: (SRBFStrategy, Levy(dim=30), 2000, "simasync", {}, "trace_srbf.npz"),
           5: (SOPStrategy, Rastrigin(dim=20), 3000, "simasync", {"ncenters": 8}, "trace_sop.npz")}[args.config]
    strategy_cls, prob, max_evals, ctrl_kind, skw, golden = cfg
    max_evals = args.evals or max_evals
    ctrl_kind = args.controller or ctrl_kind
    d = prob.dim

    out = {"config": args.config, "strategy": strategy_cls.__name__, "surrogate": args.surrogate,
           "controller": ctrl_kind, "max_evals": max_evals, "pySOT": where, "generator": args.generator}
    if args.surrogate == "gpu":
        import pysot_b200 as pb

        assert pb.HAVE_PYSOT, "pysot_b200 was imported before pySOT was importable"
        out["installed"] = pb.install()
        pb.set_candidate_generator(args.generator, seed=args.seed)
        rbf = pb.GpuRBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub, kernel=CubicKernel(), tail=LinearTail(d))
        out["isinstance_pySOT_RBFInterpolant"] = isinstance(rbf, RBFInterpolant)
    else:
        rbf = RBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub, kernel=CubicKernel(), tail=LinearTail(d))
        if args.config == 5:  # as in tools/make_strategy_traces.py: identical output, hours faster
            import pySOT.strategy.sop_strategy as m_sop
            from pysot_b200.pareto import nd_sorting

            m_sop.nd_sorting = nd_sorting
    out["isinstance_pySOT_Surrogate"] = isinstance(rbf, Surrogate)

    evals_done = {"n": 0}

    class Abort(Exception):
        pass

    abort_at = max_evals // 2 if args.checkpoint else None

    def objective(x):
        evals_done["n"] += 1
        if abort_at is not None and evals_done["n"] == abort_at:
            raise Abort()
        return prob.eval(x)

    def make_controller():
        if ctrl_kind == "serial":
            return poap.SerialController(objective)
        if ctrl_kind == "simasync":
            return poap.SimAsyncController(objective, args.workers)
        c = poap.ThreadController()
        for _ in range(args.workers):
            c.launch_worker(poap.BasicWorkerThread(c, prob.eval))
        return c

    np.random.seed(args.seed)
    slhd = SymmetricLatinHypercube(dim=d, num_pts=2 * (d + 1))
    strat = strategy_cls(max_evals=max_evals, opt_prob=prob, exp_design=slhd, surrogate=rbf, asynchronous=True, **skw)
    controller = make_controller()
    controller.strategy = strat
    t0 = time.perf_counter()
    if args.checkpoint:
        assert ctrl_kind == "serial", "--checkpoint drives the serial controller"
        fname = "live_checkpoint_%d.pysot" % os.getpid()
        os.chdir(os.environ.get("TMPDIR", "/tmp"))  # strategy.save writes temp_<fname> into the cwd (surrogate_strategy.py:177)
        ck = CheckpointController(controller, fname=fname)
        try:
            ck.run()
            raise SystemExit("the objective should have aborted the run")
        except Abort:
            pass
        n_before = strat.num_evals
        hist_before = strat.X.copy()
        abort_at = None
        del strat, controller, ck, rbf
        controller = make_controller()
        ck = CheckpointController(controller, fname=fname)
        result = ck.resume()  # dill.load -> strategy.resume() -> run (controller.py:79-93)
        strat = controller.strategy
        rbf = strat.surrogate
        os.remove(fname)
        # CheckpointController's record callback runs before the strategy's (it is attached first, controller.py:42-50),
        # so the snapshot taken at completion k holds k - 1 evaluations
        k = n_before - 1
        out["checkpoint"] = {"evals_before_abort": int(n_before),
                             "prefix_identical_after_resume": bool(np.array_equal(strat.X[:k], hist_before[:k])),
                             "surrogate_type_after_resume": type(rbf).__name__}
    else:
        result = controller.run()
    out["seconds"] = time.perf_counter() - t0

    # the reference's own bookkeeping checks (tests/test_strategies.py:13-31)
    assert strat.num_evals == max_evals and strat.pending_evals == 0 and strat.init_pending == 0
    assert strat.X.shape == (max_evals, d) and strat.fX.shape == (max_evals, 1) and strat.Xpend.shape == (0, d)
    assert len(strat.fevals) == max_evals
    assert np.all(strat.X <= prob.ub) and np.all(strat.X >= prob.lb)
    out.update(n_evals=int(strat.num_evals), best=float(result.value), rejected=int(strat.rejected_count),
               hist_sha256=hashlib.sha256(np.ascontiguousarray(strat.X).tobytes()).hexdigest()[:16])
    if args.config == 5:  # BASELINE config #5: gradients at the proposed points
        g = rbf.predict_deriv(strat.X[-8:])
        out["grad_finite"] = bool(np.all(np.isfinite(g)))
    if args.surrogate == "gpu":
        out["gpu_launches"] = int(rbf.handle.launches)
        out["last_fit_method"] = rbf.handle.last_fit_method

    gpath = os.path.join(ROOT, "tests", "golden", golden)
    deterministic = ctrl_kind in ("serial", "simasync") and not args.checkpoint and args.generator == "host" and args.seed == 0
    if deterministic and os.path.exists(gpath):
        with np.load(gpath) as z:
            if "hist_X" in z.files and int(z["max_evals"]) == max_evals and ctrl_kind == cfg[3] and args.workers == 8:
                hx, hf = z["hist_X"][:max_evals], z["hist_fX"][:max_evals]
                same = np.all(hx == strat.X, axis=1)
                out["history_identical_to_reference"] = bool(same.all() and np.array_equal(hf, strat.fX))
                out["first_difference"] = None if same.all() else int(np.argmin(same))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
