"""CPU: the oracle (oracle/rbf_oracle.py) against the fixtures generated from the unmodified reference.

tools/make_golden.py asserted bit-for-bit agreement on the build machine; here the tolerance is 1e-12 of
the result's magnitude because LAPACK / SVML kernels may differ between host CPU models.
"""
import numpy as np
import pytest

from conftest import GOLDEN, Problem, relmax
from oracle import rbf_oracle as orc

TOL = 1e-11


def split(name):
    k, t = name.split("_")[:2]
    return k, t


def test_fit_predict_deriv(fit_cases):
    assert len(fit_cases.names) == 60
    for name in fit_cases.names:
        g = fit_cases.get(name)
        kernel, tail = split(name)
        d = g["X"].shape[1]
        o = orc.OracleRBF(d, g["lb"], g["ub"], kernel, tail)
        o.add_points(g["X"], g["fX"])
        assert relmax(o.predict(g["q"]), g["pred"]) < TOL, name
        assert relmax(o.predict_deriv(g["dq"]), g["deriv"]) < TOL, name


def test_incremental_refits():
    g = np.load(GOLDEN + "/incremental.npz")
    o = orc.OracleRBF(2, np.zeros(2), np.ones(2))
    for i in range(9):
        o.add_points(g["X"][i * 100:(i + 1) * 100], g["fX"][i * 100:(i + 1) * 100])
        assert relmax(o.predict(g["Xs"]), g["step_pred"][i]) < 1e-9, i
    assert relmax(o.predict_deriv(g["Xs"]), g["inc_deriv"]) < 1e-8
    # the reference's own accuracy test (tests/test_surrogates.py:121-163)
    Xs = g["Xs"]
    f = Xs[:, 1] * np.sin(Xs[:, 0]) + Xs[:, 0] * np.cos(Xs[:, 1])
    assert np.max(np.abs(f[:, None] - o.predict(Xs))) < 1e-4


def test_capped():
    g = np.load(GOLDEN + "/capped.npz")
    o = orc.OracleRBF(3, g["lb"], g["ub"], output_transformation=orc.median_capping)
    o.add_points(g["X"], g["fX"])
    assert relmax(o.predict(g["q"]), g["pred"]) < TOL
    assert relmax(o.predict_deriv(g["q"]), g["deriv"]) < TOL


def test_merit_selection(merit_cases):
    for name in merit_cases.names:
        g = merit_cases.get(name)
        d = g["X"].shape[1]
        o = orc.OracleRBF(d, g["lb"], g["ub"], str(g["kernel"]), str(g["tail"]))
        o.add_points(g["X"], g["fX"])
        tr = {}
        with np.errstate(invalid="ignore"):
            pts = orc.weighted_distance_merit(len(g["weights"]), o, g["X"], g["fX"], g["cand"].copy(), list(g["weights"]),
                                              g.get("Xpend"), float(g["dtol"]), trace=tr)
        assert np.array_equal(tr["picked"], g["picked"]), name
        assert np.array_equal(pts, g["new_points"]), name
        assert np.array_equal(tr["dmerit0"], g["dmerit0"]), name  # cdist is plain C: bit-exact everywhere
        assert relmax(tr["fvals_raw"], g["fvals_raw"]) < TOL, name


def test_candidate_generators_seeded(candidate_cases):
    for name in candidate_cases.names:
        g = candidate_cases.get(name)
        prob = Problem(g["lb"], g["ub"], g["int_var"])
        kw = {k[3:]: (v.item() if v.ndim == 0 else v) for k, v in g.items() if k.startswith("kw_")}
        if "subset" in kw:
            kw["subset"] = list(np.atleast_1d(kw["subset"]))
        gen = dict(dycors=orc.gen_dycors, srbf=orc.gen_srbf, uniform=orc.gen_uniform)[str(g["kind"])]
        np.random.seed(int(g["seed"]))
        cand = gen(prob, g["X"], g["fX"], **kw)
        assert np.array_equal(cand, g["cand"]), name


def test_known_answer():
    """reference tests/test_auxiliary_problems.py:8-31: w = 0.25 on 1-D Ackley picks the middle of the largest gap."""
    g = np.load(GOLDEN + "/known_answer.npz")
    assert abs(g["x_uniform"][0, 0] - 10.5) < 1e-2 and abs(g["x_srbf"][0, 0] - 10.5) < 1e-2
    prob = Problem(g["lb"], g["ub"])
    o = orc.OracleRBF(1, g["lb"], g["ub"])
    o.add_points(g["X"], g["fX"])
    np.random.seed(0)
    x = orc.candidate_uniform(1, prob, o, g["X"], g["fX"], [0.25], num_cand=10000)
    assert np.array_equal(x, g["x_uniform"])


def test_unit_rescale_and_tiling():
    assert np.array_equal(orc.unit_rescale(np.array([[3.0], [3.0]])), np.ones((2, 1)))  # tests/test_utils.py:46-50
    rng = np.random.default_rng(5)
    o = orc.OracleRBF(4, np.zeros(4), np.ones(4))
    X = rng.random((40, 4))
    o.add_points(X, rng.random((40, 1)))
    cand = rng.random((1000, 4))
    fv, dm = orc.score_tiled(o, X, cand, chunk=128, threads=2)
    tr = {}
    orc.weighted_distance_merit(3, o, X, None, cand, [0.3, 0.5, 0.8], trace=tr)
    assert np.array_equal(dm, tr["dmerit0"]) and relmax(fv, tr["fvals_raw"]) < 1e-13
    assert np.array_equal(orc.select_from_scores(cand, fv, dm, [0.3, 0.5, 0.8], 3), tr["picked"])


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/pySOT"), reason="reference tree not present")
def test_oracle_matches_live_reference():
    """Build container only: oracle == live reference, bit for bit, on a fresh random case."""
    from tools import ref_import

    ref_import.load()
    from pySOT.surrogate import RBFInterpolant, TPSKernel, LinearTail

    rng = np.random.default_rng(77)
    d, n = 7, 90
    lb, ub = -np.ones(d), 2 * np.ones(d)
    X = lb + (ub - lb) * rng.random((n, d))
    fX = rng.standard_normal((n, 1))
    r = RBFInterpolant(dim=d, lb=lb, ub=ub, kernel=TPSKernel(), tail=LinearTail(d))
    r.add_points(X, fX)
    o = orc.OracleRBF(d, lb, ub, "tps", "linear")
    o.add_points(X, fX)
    q = lb + (ub - lb) * rng.random((33, d))
    assert np.array_equal(r.predict(q), o.predict(q))
    assert np.array_equal(r.predict_deriv(q), o.predict_deriv(q))


def test_strategy_trace_replay_with_the_oracle():
    """The recorded reference proposals (config #1 in full, the first calls of #2 and #5) are reproduced by the
    oracle under teacher-forced replay -- this also pins the replay logic the GPU test relies on."""
    from trace_replay import OracleBackend, replay

    ncalls, mismatches, ties, _ = replay("trace_dycors.npz", OracleBackend())
    assert ncalls > 400 and not mismatches and ties == 0
    for fname in ("trace_srbf.npz", "trace_sop.npz"):
        ncalls, mismatches, ties, err = replay(fname, OracleBackend(), max_calls=120)
        assert ncalls == 120 and not mismatches and ties == 0
        if err:  # the oracle extends its LU exactly like the reference: bit-identical to the recorded values
            assert err["pred_vs_incremental"] == 0.0 and err["deriv_vs_incremental"] == 0.0
