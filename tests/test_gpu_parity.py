"""GPU (B200): the CUDA path through the C ABI against the reference's golden vectors and the oracle.

Tolerances (BASELINE.json north_star / SURVEY 8(c)):
  predictions, gradients, coefficients-through-predictions:  max|delta| <= 1e-10 * max|reference|
  selected candidate indices: identical, except where the reference's own merits tie within 1e-12
  distances in exact mode: bit-identical to scipy's cdist
"""
import numpy as np
import pytest

from conftest import GOLDEN, Problem, relmax

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def pb():
    import pysot_b200

    return pysot_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import rbf_oracle

    return rbf_oracle


KERNELS = {"cubic": "CubicKernel", "tps": "TPSKernel", "linear": "LinearKernel"}
TAILS = {"linear": "LinearTail", "constant": "ConstantTail"}


def make_gpu(pb, d, lb, ub, kernel="cubic", tail="linear", **kw):
    return pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=getattr(pb, KERNELS[kernel])(),
                                tail=getattr(pb, TAILS[tail])(d), **kw)


# ------------------------------------------------------------------------------------------------ building blocks
def test_dmma_gemm_against_numpy(pb):
    from pysot_b200 import _lib

    h = _lib.Handle(0)
    rng = np.random.default_rng(0)
    for M, Nc, K in [(128, 128, 16), (300, 260, 32), (77, 1, 32), (513, 131, 128), (1000, 2, 64)]:
        A, B, C = rng.standard_normal((M, K)), rng.standard_normal((K, Nc)), rng.standard_normal((M, Nc))
        got, _ = h.dbg_gemm(C, A, B)
        ref = C - A @ B
        assert relmax(got, ref) < 1e-13, (M, Nc, K)


def lu_residual(LU, ipiv, A):
    N = A.shape[0]
    L = np.tril(LU, -1) + np.eye(N)
    U = np.triu(LU)
    PA = A.copy()
    for i in range(N):
        p = ipiv[i]
        if p != i:
            PA[[i, p]] = PA[[p, i]]
    return np.max(np.abs(L @ U - PA)) / np.max(np.abs(A)), L


@pytest.mark.parametrize("N", [1, 5, 31, 32, 33, 127, 128, 129, 200, 517, 1100])
def test_pivoted_lu_against_numpy(pb, N):
    import scipy.linalg as sla
    from pysot_b200 import _lib

    h = _lib.Handle(0)
    rng = np.random.default_rng(N)
    A = rng.standard_normal((N, N))
    if N > 8:
        A[:3, :3] = 0.0  # zero leading block like the saddle-point system: pivoting is mandatory
    b = rng.standard_normal((N, 1))
    LU, ipiv, info = h.dbg_lu(np.hstack((A, b)), nrhs=1)
    assert info == 0
    res, L = lu_residual(LU[:, :N], ipiv, A)
    assert res < 1e-13, res
    assert np.max(np.abs(L)) <= 1.0 + 1e-15  # partial pivoting: multipliers bounded by one
    _, piv_ref = sla.lu_factor(A)
    assert np.array_equal(ipiv, piv_ref)  # same pivot sequence as LAPACK dgetrf
    y = LU[:, N]
    x = sla.solve_triangular(np.triu(LU[:, :N]), y)
    assert relmax(x, np.linalg.solve(A, b.ravel())) < 1e-9


def test_lu_reports_exact_singularity(pb):
    from pysot_b200 import _lib

    A = np.ones((40, 40))
    _, _, info = _lib.Handle(0).dbg_lu(A)
    assert info == 2


# ------------------------------------------------------------------------------------------------ fit / predict / deriv
def test_fit_predict_deriv_golden(pb, fit_cases):
    worst = {}
    for name in fit_cases.names:
        g = fit_cases.get(name)
        kernel, tail = name.split("_")[:2]
        d = g["X"].shape[1]
        s = make_gpu(pb, d, g["lb"], g["ub"], kernel, tail)
        s.add_points(g["X"], g["fX"])
        ep = relmax(s.predict(g["q"]), g["pred"])
        ed = relmax(s.predict_deriv(g["dq"]), g["deriv"])
        worst[name] = (ep, ed)
        assert s.c.shape == g["c"].shape and s.updated
    bad = {k: v for k, v in worst.items() if not (v[0] < TOL and v[1] < TOL)}
    assert not bad, bad


def test_fit_exact_mode_golden(pb, fit_cases):
    for name in fit_cases.names[::7]:
        g = fit_cases.get(name)
        kernel, tail = name.split("_")[:2]
        s = make_gpu(pb, g["X"].shape[1], g["lb"], g["ub"], kernel, tail, exact_dist=True)
        s.add_points(g["X"], g["fX"])
        assert relmax(s.predict(g["q"]), g["pred"]) < TOL, name
        assert relmax(s.predict_deriv(g["dq"]), g["deriv"]) < TOL, name


def test_reference_test_rbf_scenario(pb):
    """tests/test_surrogates.py:121-163 re-expressed: accuracy against the analytic function, reset, nine
    batches with forced refits (the reference extends its LU there; the GPU refits), same bounds again."""
    g = np.load(GOLDEN + "/incremental.npz")
    X, fX, Xs = g["X"], g["fX"], g["Xs"]
    f = (Xs[:, 1] * np.sin(Xs[:, 0]) + Xs[:, 0] * np.cos(Xs[:, 1]))[:, None]
    df = np.stack((Xs[:, 1] * np.cos(Xs[:, 0]) + np.cos(Xs[:, 1]), np.sin(Xs[:, 0]) - Xs[:, 0] * np.sin(Xs[:, 1])), 1)
    s = pb.GpuRBFInterpolant(dim=2, lb=np.zeros(2), ub=np.ones(2), eta=1e-6)
    assert isinstance(s, pb.Surrogate)
    s.add_points(X, fX)
    assert np.max(np.abs(f - s.predict(Xs))) < 1e-4
    assert np.linalg.norm(df - s.predict_deriv(Xs)) < 1e-2
    assert relmax(s.predict(Xs), g["full_pred"]) < TOL and relmax(s.predict_deriv(Xs), g["full_deriv"]) < TOL
    x0 = X[0]
    df0 = np.array([[x0[1] * np.cos(x0[0]) + np.cos(x0[1]), np.sin(x0[0]) - x0[0] * np.sin(x0[1])]])
    assert np.linalg.norm(df0 - s.predict_deriv(x0)) < 1e-1
    s.reset()
    assert s.num_pts == 0 and s.dim == 2
    for i in range(9):
        s.add_points(X[i * 100:(i + 1) * 100], fX[i * 100:(i + 1) * 100])
        assert relmax(s.predict(Xs), g["step_pred"][i]) < 1e-9, i  # reference's incremental path: its own noise
    assert np.max(np.abs(f - s.predict(Xs))) < 1e-4
    assert np.linalg.norm(df - s.predict_deriv(Xs)) < 1e-2


def test_capped_and_unit_box(pb):
    """tests/test_surrogates.py:229-280: median capping == pre-capped values; identical surrogates agree; reset."""
    g = np.load(GOLDEN + "/capped.npz")
    s1 = pb.GpuRBFInterpolant(dim=3, lb=g["lb"], ub=g["ub"], output_transformation=pb.median_capping)
    s1.add_points(g["X"], g["fX"])
    s2 = pb.GpuRBFInterpolant(dim=3, lb=g["lb"], ub=g["ub"])
    s2.add_points(g["X"], pb.median_capping(g["fX"]))
    assert np.max(np.abs(s1.predict(g["q"]) - s2.predict(g["q"]))) < 1e-10
    assert np.max(np.abs(s1.predict_deriv(g["q"]) - s2.predict_deriv(g["q"]))) < 1e-10
    assert relmax(s1.predict(g["q"]), g["pred"]) < TOL and relmax(s1.predict_deriv(g["q"]), g["deriv"]) < TOL
    s1.reset()
    assert s1.X.shape == (0, 3) and s1.fX.shape == (0, 1)


def test_shapes_errors_and_pickle(pb):
    import dill

    rng = np.random.default_rng(1)
    s = pb.GpuRBFInterpolant(dim=4, lb=np.zeros(4), ub=2 * np.ones(4))
    X = 2 * rng.random((30, 4))
    s.add_points(X, np.sin(X.sum(1)))
    assert s.predict(X[0]).shape == (1, 1) and s.predict(X[:7]).shape == (7, 1)
    assert s.predict_deriv(X[0]).shape == (1, 4) and s.predict_deriv(X[:7]).shape == (7, 4)
    with pytest.raises(ValueError):
        s.predict(np.zeros((3, 5)))
    with pytest.raises(ValueError):
        s.predict_deriv(np.zeros((3, 5)))
    p0 = s.predict(X[:5] + 0.01)
    t = dill.loads(dill.dumps(s))  # CheckpointController round trip: handle dropped, refit on demand
    assert t._handle is None and t.c is None
    assert np.max(np.abs(t.predict(X[:5] + 0.01) - p0)) < 1e-12
    few = pb.GpuRBFInterpolant(dim=4, lb=np.zeros(4), ub=np.ones(4))
    few.add_points(rng.random((3, 4)), rng.random(3))
    with pytest.raises(AssertionError):  # rbf.py:111
        few.predict(np.zeros(4))


def test_interpolates_and_duplicate_centers(pb, orc):
    """eta-regularised interpolation reproduces the data; duplicate points (optimiser-like clouds) stay finite."""
    rng = np.random.default_rng(2)
    d, n = 6, 150
    X = rng.random((n, d))
    X[10] = X[3]  # exact duplicate
    X[11] = X[4] + 1e-9
    fX = np.cos(X.sum(1))[:, None]
    for kernel in ("cubic", "tps"):
        s = make_gpu(pb, d, np.zeros(d), np.ones(d), kernel, "linear")
        s.add_points(X, fX)
        o = orc.OracleRBF(d, np.zeros(d), np.ones(d), kernel, "linear")
        o.add_points(X, fX)
        q = rng.random((50, d))
        assert relmax(s.predict(q), o.predict(q)) < 1e-9, kernel  # cond(A) ~ 1/eta here: reference noise is 1e-10
        assert np.max(np.abs(s.predict(X) - fX)) < 1e-3


# ------------------------------------------------------------------------------------------------ acquisition
def check_selection(name, idx, g, orc, fv, dm):
    ref = g["picked"]
    if np.array_equal(idx, ref):
        return
    # allowed only where the reference's own merits tie within 1e-12 (north_star)
    o = orc.OracleRBF(g["X"].shape[1], g["lb"], g["ub"], str(g["kernel"]), str(g["tail"]))
    o.add_points(g["X"], g["fX"])
    tr = {}
    with np.errstate(invalid="ignore"):
        orc.weighted_distance_merit(len(ref), o, g["X"], g["fX"], g["cand"].copy(), list(g["weights"]), g.get("Xpend"),
                                    float(g["dtol"]), trace=tr)
    for i, (a, b) in enumerate(zip(idx, ref)):
        if a != b:
            m = tr["merit"][i]
            assert abs(m[a, 0] - m[b, 0]) <= 1e-12, (name, i, int(a), int(b), m[a, 0], m[b, 0])
            break  # later picks legitimately diverge after a tie


@pytest.mark.parametrize("mode", ["mma", "direct", "exact"])
def test_merit_selection_golden(pb, orc, merit_cases, mode):
    """mma: norm expansion with DMMA dot products (default); direct: FMA differences; exact: cdist-identical."""
    from pysot_b200 import _lib

    exact = mode == "exact"
    for name in merit_cases.names:
        g = merit_cases.get(name)
        d = g["X"].shape[1]
        s = make_gpu(pb, d, g["lb"], g["ub"], str(g["kernel"]), str(g["tail"]), exact_dist=exact)
        s.handle.set_option(_lib.OPT_MMA_DIST, 1 if mode == "mma" else 0)
        s.add_points(g["X"], g["fX"])
        p = len(g["weights"])
        idx, merit, fv, dm, st = pb.merit_select_indices(p, s, g["X"], g["cand"], list(g["weights"]), g.get("Xpend"),
                                                         float(g["dtol"]), want_scores=True)
        assert relmax(fv, g["fvals_raw"].ravel()) < TOL, name
        if exact and np.all(g["ub"] - g["lb"] == 1.0) and np.all(g["lb"] == 0.0):
            assert np.array_equal(dm, g["dmerit0"].ravel()), name  # unit box: one distance, bit-identical to cdist
        elif exact and np.any(g["ub"] - g["lb"] != (g["ub"] - g["lb"])[0]):
            assert np.array_equal(dm, g["dmerit0"].ravel()), name  # anisotropic box: original-space pass, exact
        else:
            assert relmax(dm, g["dmerit0"].ravel()) < 1e-12, name
            assert np.array_equal(dm == 0.0, g["dmerit0"].ravel() == 0.0), name  # coincident points stay exactly 0
        assert st[0] == fv.min() and st[1] == fv.max() and st[2] == dm.min() and st[3] == dm.max(), name
        check_selection(name, idx, g, orc, fv, dm)
        pts = pb.weighted_distance_merit(p, s, g["X"], g["fX"], g["cand"], list(g["weights"]), g.get("Xpend"),
                                         float(g["dtol"]))
        assert pts.shape == (p, d)
        if np.array_equal(idx, g["picked"]):
            assert np.array_equal(pts, g["new_points"]), name
            ok = np.isfinite(g["picked_merit"])
            assert np.allclose(merit[ok], g["picked_merit"][ok], rtol=0, atol=1e-12), name


def test_merit_without_aliasing(pb, orc):
    """extra_points case (surrogate_strategy.py:243-247): the surrogate holds points that X does not."""
    rng = np.random.default_rng(3)
    d = 5
    lb, ub = np.zeros(d), 3 * np.ones(d)
    Xall = 3 * rng.random((60, d))
    fall = np.sin(Xall.sum(1))[:, None]
    X, fX = Xall[10:], fall[10:]
    cand = 3 * rng.random((900, d))
    s = make_gpu(pb, d, lb, ub)
    s.add_points(Xall, fall)
    o = orc.OracleRBF(d, lb, ub)
    o.add_points(Xall, fall)
    tr = {}
    ref = orc.weighted_distance_merit(3, o, X, fX, cand.copy(), [0.3, 0.5, 0.8], None, 1e-3, trace=tr)
    idx, _, fv, dm, _ = pb.merit_select_indices(3, s, X, cand, [0.3, 0.5, 0.8], None, 1e-3, want_scores=True)
    assert relmax(dm, tr["dmerit0"].ravel()) < 1e-13 and relmax(fv, tr["fvals_raw"].ravel()) < TOL
    assert np.array_equal(idx, tr["picked"])
    assert np.array_equal(pb.weighted_distance_merit(3, s, X, fX, cand, [0.3, 0.5, 0.8]), ref)


def test_candidate_functions_seeded(pb, candidate_cases):
    """Seeded end-to-end parity of candidate_dycors / candidate_srbf / candidate_uniform with the reference."""
    for name in candidate_cases.names:
        g = candidate_cases.get(name)
        prob = Problem(g["lb"], g["ub"], g["int_var"])
        d = prob.dim
        kw = {k[3:]: (v.item() if v.ndim == 0 else v) for k, v in g.items() if k.startswith("kw_")}
        if "subset" in kw:
            kw["subset"] = list(np.atleast_1d(kw["subset"]))
        s = pb.GpuRBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub)
        s.add_points(g["X"], g["fX"])
        fn = dict(dycors=pb.candidate_dycors, srbf=pb.candidate_srbf, uniform=pb.candidate_uniform)[str(g["kind"])]
        np.random.seed(int(g["seed"]))
        pts = fn(num_pts=len(g["weights"]), opt_prob=prob, surrogate=s, X=g["X"], fX=g["fX"], weights=list(g["weights"]),
                 Xpend=g["Xpend"], **kw)
        assert np.array_equal(pts, g["new_points"]), name


def test_reference_known_answer(pb):
    """tests/test_auxiliary_problems.py:8-31 with the RBF surrogate: w = 0.25 -> 10.50 (+- 1e-2)."""
    g = np.load(GOLDEN + "/known_answer.npz")
    prob = Problem(g["lb"], g["ub"])
    s = pb.GpuRBFInterpolant(dim=1, lb=g["lb"], ub=g["ub"])
    s.add_points(g["X"], g["fX"])
    np.random.seed(0)
    xu = pb.candidate_uniform(num_pts=1, X=g["X"], Xpend=None, fX=g["fX"], num_cand=10000, surrogate=s, opt_prob=prob,
                              weights=[0.25])
    xs = pb.candidate_srbf(num_pts=1, X=g["X"], Xpend=None, fX=g["fX"], num_cand=10000, surrogate=s, opt_prob=prob,
                           weights=[0.25], sampling_radius=0.5)
    assert np.isclose(xu, 10.50, atol=1e-2) and np.isclose(xs, 10.50, atol=1e-2)
    assert np.array_equal(xu, g["x_uniform"]) and np.array_equal(xs, g["x_srbf"])


# ------------------------------------------------------------------------------------------------ larger sizes
def test_large_scoring_against_tiled_oracle(pb, orc):
    """config #4 shape at a size the oracle finishes in seconds: d=50, 2 000 centers, 40 000 candidates,
    chunked host pipeline (several chunks) vs the tiled oracle; then device-resident input vs host input."""
    import torch

    rng = np.random.default_rng(4)
    d, n, m = 50, 2000, 120000  # > 2 waves of the score kernel (148 SMs x 3 CTAs x 128 rows): several chunks
    lb, ub = np.zeros(d), np.ones(d)
    X = rng.random((n, d))
    c = rng.standard_normal((n + d + 1, 1))
    xbest = X[0]
    cand = np.where(rng.random((m, d)) < 0.4, np.clip(xbest + 0.2 * rng.standard_normal((m, d)), 0, 1), xbest)
    Xpend = rng.random((7, d))
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, exact_dist=True)
    s.add_points(X, np.zeros(n))
    s.set_coefficients(c)
    s.handle.set_option(2, 8192)  # B200RBF_OPT_CHUNK_ROWS: rounded up to one full wave per chunk -> 3 chunks
    o = orc.OracleRBF(d, lb, ub)
    o.add_points(X, np.zeros(n))
    o.c, o.updated = c, True
    fv_ref, dm_ref = orc.score_tiled(o, X, cand, Xpend, chunk=4096, threads=4)
    w = [0.3, 0.5, 0.8, 0.95]
    idx, _, fv, dm, _ = pb.merit_select_indices(4, s, X, cand, w, Xpend, want_scores=True)
    assert relmax(fv, fv_ref.ravel()) < TOL
    assert np.array_equal(dm, dm_ref.ravel())
    assert np.array_equal(idx, orc.select_from_scores(cand, fv_ref, dm_ref, w, 4))
    # same rows handed over as a CUDA tensor (no staging): identical scores and picks
    ct = torch.from_numpy(cand).cuda()
    fv2, dm2, _ = s.handle.score(ct, lb, ub, np.vstack((X, Xpend)), True, want_scores=True, m=m)
    idx2, _ = s.handle.select(w, 4, 1e-3)
    assert np.array_equal(fv2, fv) and np.array_equal(dm2, dm) and np.array_equal(idx2, idx)
    out = s.predict_device(ct[:1000])
    assert np.array_equal(out.cpu().numpy().ravel(), fv[:1000])
    # weighted_distance_merit on the tensor: the rows come back with the indices (b200rbf_select_rows)
    pts = pb.weighted_distance_merit(4, s, X, None, ct, w, Xpend)
    assert pts.shape == (4, d) and np.array_equal(pts, cand[idx])
    assert np.array_equal(pb.weighted_distance_merit(4, s, X, None, cand, w, Xpend), cand[idx])


def test_fit_2000_centers_tps(pb, orc):
    """config #3 shape (TPS + linear tail, d=10) at N = 2 000 against scipy's dgetrf path."""
    rng = np.random.default_rng(0)
    d, n = 10, 2000
    X = rng.random((n, d))
    fX = (np.sin(3 * X.sum(1)) + X[:, 0] ** 2)[:, None]
    s = make_gpu(pb, d, np.zeros(d), np.ones(d), "tps", "linear")
    s.add_points(X, fX)
    o = orc.OracleRBF(d, np.zeros(d), np.ones(d), "tps", "linear")
    o.add_points(X, fX)
    q = rng.random((500, d))
    assert relmax(s.predict(q), o.predict(q)) < TOL
    assert relmax(s.predict_deriv(q[:20]), o.predict_deriv(q[:20])) < TOL
    # property at any size: the fitted model satisfies the saddle-point equations
    assert np.max(np.abs(s.predict(X) + s.eta * s.c[d + 1:] - fX)) < 1e-8 * np.max(np.abs(s.c))
    P = np.hstack((np.ones((n, 1)), X))
    assert np.max(np.abs(P.T @ s.c[d + 1:])) < 1e-7 * np.max(np.abs(s.c))


def test_properties_at_scale(pb):
    """Size-independent checks at a size no CPU oracle is run for (d=50, 20 000 centers, 2^18 candidates)."""
    rng = np.random.default_rng(5)
    d, n, m = 50, 20000, 1 << 18
    X = rng.random((n, d))
    c = rng.standard_normal((n + d + 1, 1))
    s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d))
    s.add_points(X, np.zeros(n))
    s.set_coefficients(c)
    cand = rng.random((m, d))
    cand[12345] = X[777]
    idx, merit, fv, dm, st = pb.merit_select_indices(4, s, X, cand, [0.3, 0.5, 0.8, 0.95], None, want_scores=True)
    assert dm[12345] == 0.0 and np.all(dm >= 0) and np.all(np.isfinite(fv))
    assert st[0] == fv.min() and st[1] == fv.max() and st[2] == dm.min() and st[3] == dm.max()
    assert len(set(idx.tolist())) == 4 and 12345 not in idx  # picked rows are distinct and never within dtol
    # linearity in the coefficients: s(2c) = 2 s(c)
    s.set_coefficients(2 * c)
    fv2 = s.predict(cand[:4096]).ravel()
    assert relmax(fv2, 2 * fv[:4096]) < 1e-13
    # permutation of the candidates permutes the scores
    perm = rng.permutation(4096)
    assert np.array_equal(s.predict(cand[:4096][perm]).ravel(), fv2[perm])


@pytest.mark.parametrize("d", [66, 101])
def test_dimensions_above_64_use_the_generic_kernel(pb, orc, d):
    """padded dim > 64: candidate coordinates in shared memory instead of registers (score_generic.cu)."""
    rng = np.random.default_rng(d)
    n, m = d + 40, 3000
    lb, ub = -np.ones(d), np.ones(d) * 1.5
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(X.sum(1) / 5)[:, None]
    Xpend = lb + (ub - lb) * rng.random((2, d))
    cand = lb + (ub - lb) * rng.random((m, d))
    for kernel in ("cubic", "tps"):
        s = make_gpu(pb, d, lb, ub, kernel, "linear", exact_dist=True)
        s.add_points(X, fX)
        o = orc.OracleRBF(d, lb, ub, kernel, "linear")
        o.add_points(X, fX)
        tr = {}
        orc.weighted_distance_merit(3, o, X, fX, cand.copy(), [0.3, 0.5, 0.8], Xpend, 1e-3, trace=tr)
        idx, _, fv, dm, _ = pb.merit_select_indices(3, s, X, cand, [0.3, 0.5, 0.8], Xpend, want_scores=True)
        assert relmax(fv, tr["fvals_raw"].ravel()) < TOL and relmax(dm, tr["dmerit0"].ravel()) < 1e-13
        assert np.array_equal(idx, tr["picked"])
        assert relmax(s.predict_deriv(cand[:5]), o.predict_deriv(cand[:5])) < TOL
    big = pb.GpuRBFInterpolant(dim=500, lb=np.zeros(500), ub=np.ones(500))
    big.add_points(rng.random((501, 500)), np.zeros(501))
    with pytest.raises(ValueError):  # beyond the generic kernel's shared-memory budget: refused, never a CPU fallback
        big.predict(np.zeros(500))


def test_mma_and_direct_kernels_agree_on_clustered_points(pb, orc):
    """Optimiser-like cloud: centers and candidates clustered around one point, many NEAR pairs (the DMMA kernel's
    direct-recompute path), duplicates included; both kernels against the oracle."""
    from pysot_b200 import _lib

    rng = np.random.default_rng(11)
    d, n, m = 30, 400, 6000
    lb, ub = np.zeros(d), np.ones(d)
    x0 = rng.random(d)
    X = np.clip(x0 + 10.0 ** rng.uniform(-6, -0.5, (n, 1)) * rng.standard_normal((n, d)), 0, 1)
    fX = ((X - x0) ** 2).sum(1, keepdims=True)
    cand = np.clip(x0 + 10.0 ** rng.uniform(-7, -0.5, (m, 1)) * rng.standard_normal((m, d)), 0, 1)
    cand[:50] = X[:50]
    o = orc.OracleRBF(d, lb, ub)
    o.add_points(X, fX)
    tr = {}
    orc.weighted_distance_merit(4, o, X, fX, cand.copy(), [0.3, 0.5, 0.8, 0.95], None, 1e-3, trace=tr)
    res = {}
    for mode in ("mma", "direct"):
        s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
        s.handle.set_option(_lib.OPT_MMA_DIST, 1 if mode == "mma" else 0)
        s.add_points(X, fX)
        s.set_coefficients(o.c)  # same coefficients: isolates the evaluation kernels from the (ill-conditioned) fit
        idx, _, fv, dm, _ = pb.merit_select_indices(4, s, X, cand, [0.3, 0.5, 0.8, 0.95], None, want_scores=True)
        res[mode] = (idx, fv, dm)
        assert relmax(fv, tr["fvals_raw"].ravel()) < TOL, mode
        assert np.all(dm[:50] == 0.0) and relmax(dm, tr["dmerit0"].ravel()) < 1e-12, mode
        rel = np.abs(dm - tr["dmerit0"].ravel())[50:] / tr["dmerit0"].ravel()[50:]
        assert rel.max() < 1e-9, (mode, rel.max())  # near pairs keep RELATIVE accuracy (direct recomputation)
        assert np.array_equal(idx, tr["picked"]), mode


@pytest.mark.parametrize("kernel,tail", [("cubic", "linear"), ("tps", "linear"), ("linear", "constant")])
@pytest.mark.parametrize("d", [5, 6, 7, 8, 49, 52])
def test_score_and_gradient_kernels_over_dimension_classes(pb, orc, kernel, tail, d):
    """d mod 4 decides the DMMA kernel's variant (norms inside the MMA for 1, 2; separate for 3, 0); each kernel has its
    own epilogue (rsqrt-based square root, table-driven half-log) and its own gradient scalar.  Same coefficients on
    both sides, 5000 candidates / queries (tile kernels), against the oracle."""
    rng = np.random.default_rng(100 + d)
    n, m = 333, 5000
    lb, ub = -1.0 - rng.random(d), 1.0 + rng.random(d)
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(X.sum(1))[:, None]
    o = orc.OracleRBF(d, lb, ub, kernel, tail)
    o.add_points(X, fX)
    o.fit()
    s = make_gpu(pb, d, lb, ub, kernel, tail)
    s.add_points(X, fX)
    s.set_coefficients(o.c)
    q = lb + (ub - lb) * rng.random((m, d))
    q[:20] = X[:20]  # coincident points: r = 0 in both the value and the gradient path
    assert relmax(s.predict(q), o.predict(q)) < 1e-12
    assert relmax(s.predict_deriv(q), o.predict_deriv(q)) < 1e-11
    tr = {}
    orc.weighted_distance_merit(3, o, X, fX, q.copy(), [0.3, 0.5, 0.95], None, 1e-3, trace=tr)
    idx, _, fv, dm, _ = pb.merit_select_indices(3, s, X, q, [0.3, 0.5, 0.95], None, want_scores=True)
    assert relmax(fv, tr["fvals_raw"].ravel()) < 1e-12 and relmax(dm, tr["dmerit0"].ravel()) < 1e-12
    assert np.all(dm[:20] == 0.0) and np.array_equal(idx, tr["picked"])


@pytest.mark.parametrize("kernel,tail", [("cubic", "linear"), ("tps", "linear"), ("linear", "constant")])
@pytest.mark.parametrize("d", [10, 20])
def test_call_sizes_around_the_center_split_on_a_clustered_cloud(pb, orc, kernel, tail, d):
    """The DMMA score kernel lets 4 (2) warps share a row group when the call has few rows (up to 2368 / 4736 rows on
    148 SMs: at most one CTA per SM) and recomputes only tiny pairs and near pairs that lower the running minimum.  An optimiser-like cloud
    (eight tight clusters, radii 1e-6 .. 0.3, duplicates among the candidates) at call sizes on both sides of every
    boundary, AUG (d = 10) and plain (d = 20) variants, each kernel: values, distances (relative accuracy per
    candidate) and the pick against the oracle."""
    rng = np.random.default_rng(7 * d + len(kernel))
    n, mmax = 700, 14300
    lb, ub = -2.0 * np.ones(d), 3.0 * np.ones(d)
    cen = lb + (ub - lb) * rng.random((8, d))

    def cloud(k, lo):
        which = rng.integers(0, 8, k)
        rad = 10.0 ** rng.uniform(lo, -0.5, (k, 1)) * (ub - lb)
        return np.clip(cen[which] + rad * rng.standard_normal((k, d)) / np.sqrt(d), lb, ub)

    X = cloud(n, -6)
    fX = np.sin(X.sum(1))[:, None]
    cand = cloud(mmax, -7)
    o = orc.OracleRBF(d, lb, ub, kernel, tail)
    o.add_points(X, fX)
    o.c = rng.standard_normal((n + o.ntail, 1))  # evaluation only: no ill-conditioned fit
    o.updated = True
    s = make_gpu(pb, d, lb, ub, kernel, tail)
    s.add_points(X, fX)
    s.set_coefficients(o.c)
    for m in (1, 3, 16, 17, 100, 2000, 2368, 2369, 4736, 4737, 7105, 14209):
        q = cand[:m].copy()
        k = min(m, 5)
        q[:k] = X[:k]  # coincident with centers: r = 0 exactly
        assert relmax(s.predict(q), o.predict(q)) < TOL, m
        tr = {}
        orc.weighted_distance_merit(1, o, X, fX, q.copy(), [0.5], None, 1e-3, trace=tr)
        idx, _, fv, dm, _ = pb.merit_select_indices(1, s, X, q, [0.5], None, want_scores=True)
        ref_d = tr["dmerit0"].ravel()
        assert relmax(fv, tr["fvals_raw"].ravel()) < TOL, m
        assert np.all(dm[:k] == 0.0), m
        if m > k:
            rel = np.abs(dm - ref_d)[k:] / ref_d[k:]
            assert rel.max() < 1e-9, (m, rel.max())
        assert relmax(dm, ref_d) < 1e-12, m
        assert np.array_equal(idx, tr["picked"]), m


# n around the 32 / 128 / 512 blocking of the two-level Cholesky (padding rows, one-block case, last partial panel of
# an outer panel, first column right of an outer panel)
@pytest.mark.parametrize("kernel,d,n", [("cubic", 3, 150), ("tps", 10, 1000), ("cubic", 30, 1500), ("tps", 2, 700),
                                        ("cubic", 4, 128), ("tps", 4, 129), ("cubic", 5, 161), ("tps", 3, 256),
                                        ("cubic", 3, 257), ("tps", 6, 511), ("cubic", 6, 513), ("tps", 5, 640),
                                        ("cubic", 2, 1025)])
def test_cholesky_fit_matches_lu_and_oracle(pb, orc, kernel, d, n):
    """The null-space Cholesky solver (fit_chol.cu) against the pivoted LU and the reference's dgetrf path."""
    from pysot_b200 import _lib

    rng = np.random.default_rng(n + d)
    lb, ub = -np.ones(d), 2 * np.ones(d)
    X = lb + (ub - lb) * rng.random((n, d))
    X[n // 2:] = X[0] + 0.05 * rng.standard_normal((n - n // 2, d))  # half the points clustered (still inside the box)
    X = np.clip(X, lb, ub)
    fX = (np.sin(X.sum(1)) + X[:, 0] ** 2)[:, None]
    q = lb + (ub - lb) * rng.random((300, d))
    o = orc.OracleRBF(d, lb, ub, kernel, "linear")
    o.add_points(X, fX)
    ref_p, ref_g = o.predict(q), o.predict_deriv(q[:10])
    out = {}
    for method in (1, 2):
        s = make_gpu(pb, d, lb, ub, kernel, "linear")
        s.handle.set_option(_lib.OPT_FIT_METHOD, method)
        s.add_points(X, fX)
        out[method] = (s.predict(q), s.predict_deriv(q[:10]), s.c.copy())
        assert s.handle.last_fit_method == ("lu" if method == 1 else "cholesky")
        assert relmax(out[method][0], ref_p) < TOL, (method, relmax(out[method][0], ref_p))
        assert relmax(out[method][1], ref_g) < 1e-9, (method, relmax(out[method][1], ref_g))
    P = np.hstack((np.ones((n, 1)), (X - lb) / (ub - lb)))
    c = out[2][2]
    assert np.max(np.abs(P.T @ c[d + 1:])) < 1e-9 * np.max(np.abs(c))  # the side condition P^T c = 0
    # linear kernel / constant tail are not conditionally positive definite of order 2: explicit Cholesky is refused
    s = make_gpu(pb, d, lb, ub, "linear", "constant")
    s.handle.set_option(_lib.OPT_FIT_METHOD, 2)
    s.add_points(X, fX)
    with pytest.raises(ValueError):
        s.predict(q)


def test_lu_lookahead_and_gemm_variants_give_identical_coefficients(pb):
    """The optional LU look-ahead stream and the three DMMA GEMM variants reorder launches / tiles, not arithmetic."""
    from pysot_b200 import _lib

    rng = np.random.default_rng(9)
    d, n = 6, 900
    X = rng.random((n, d))
    fX = np.cos(2 * X.sum(1))[:, None]
    ref = None
    for look, variant, method in [(0, 3, 1), (1, 0, 1), (0, 2, 1), (0, 4, 1), (0, 3, 2), (0, 4, 2)]:
        s = make_gpu(pb, d, np.zeros(d), np.ones(d), "cubic", "linear")
        s.handle.set_option(_lib.OPT_LU_LOOKAHEAD, look)
        s.handle.set_option(_lib.OPT_GEMM_STAGES, variant)
        s.handle.set_option(_lib.OPT_FIT_METHOD, method)
        s.add_points(X, fX)
        s._fit()
        if method == 1:
            ref = s.c.copy() if ref is None else ref
            assert np.array_equal(s.c, ref), (look, variant)  # same products, same order within each accumulator
        else:
            assert relmax(s.c, ref) < 1e-7  # different algorithm: coefficients agree to cond * eps


@pytest.mark.parametrize("kernel,tail,d", [("cubic", "linear", 10), ("tps", "linear", 50), ("linear", "constant", 3),
                                           ("linear", "linear", 7), ("cubic", "linear", 64)])
def test_predict_deriv_many_queries_uses_the_tile_kernel(pb, orc, kernel, tail, d):
    """>= 4096 query points take the tile-streaming gradient kernel (deriv_tile.cuh); fewer take the slab kernel.
    Both against the oracle's Python loop (rbf.py:204-214), including queries that coincide with centers."""
    rng = np.random.default_rng(d)
    n = 2 * d + 150
    lb, ub = rng.uniform(-3, 0, d), rng.uniform(1, 4, d)
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(((X - lb) / (ub - lb)).sum(1))[:, None]
    s = make_gpu(pb, d, lb, ub, kernel, tail)
    s.add_points(X, fX)
    o = orc.OracleRBF(d, lb, ub, kernel, tail)
    o.add_points(X, fX)
    q = lb + (ub - lb) * rng.random((4500, d))
    q[:20] = X[:20]
    big = s.predict_deriv(q)
    small = s.predict_deriv(q[:300])
    ref = o.predict_deriv(q[:300])
    assert big.shape == (4500, d)
    assert relmax(small, ref) < TOL and relmax(big[:300], ref) < TOL
    ref_tail = o.predict_deriv(q[-50:])
    assert relmax(big[-50:], ref_tail) < TOL


@pytest.mark.parametrize("d", [150, 170, 190])
def test_fit_with_large_tail_dimension(pb, orc, d):
    """the Cholesky fit's t x t helper kernels keep a (t rounded to 16)^2 matrix in shared memory, which
    exceeds the 227 KB opt-in limit for d >= 160; those dimensions must take the pivoted LU instead of failing."""
    rng = np.random.default_rng(d)
    n = 2 * (d + 1) + 60
    lb, ub = np.zeros(d), np.ones(d)
    X = rng.random((n, d))
    fX = np.sin(X.sum(1) / 7)[:, None]
    s = make_gpu(pb, d, lb, ub, "cubic", "linear")
    s.add_points(X, fX)
    o = orc.OracleRBF(d, lb, ub, "cubic", "linear")
    o.add_points(X, fX)
    q = rng.random((64, d))
    assert relmax(s.predict(q), o.predict(q)) < TOL
    assert s.handle.last_fit_method == ("cholesky" if d < 160 else "lu")
    assert relmax(s.predict_deriv(q[:4]), o.predict_deriv(q[:4])) < TOL


@pytest.mark.parametrize("kernel,d,n", [("tps", 10, 3000), ("cubic", 20, 2500), ("cubic", 50, 1200), ("tps", 9, 700), ("cubic", 8, 900),
                                        ("tps", 31, 600), ("cubic", 1, 300), ("tps", 2, 300)])
def test_gradient_on_the_tensor_path_against_oracle(pb, orc, kernel, d, n):
    """deriv_mma_kernel (>= 4096 queries, cubic / TPS): grad = u sum_j s_j - S C as two chained DMMA GEMMs.  Queries on
    a clustered, optimiser-like cloud: some coincide with centers, some sit 1e-7 .. 1e-3 away (the near-pair path), the
    rest are spread over the box; every dimension class of the fragments (d mod 8) appears in the parametrisation."""
    rng = np.random.default_rng(1000 * d + n)
    lb, ub = rng.uniform(-3, 0, d), rng.uniform(1, 4, d)
    X = lb + (ub - lb) * rng.random((n, d))
    X[n // 2:] = np.clip(X[7] + 0.03 * (ub - lb) * rng.standard_normal((n - n // 2, d)), lb, ub)  # a tight cluster
    fX = np.sin(((X - lb) / (ub - lb)).sum(1) * 2)[:, None]
    s = make_gpu(pb, d, lb, ub, kernel, "linear")
    s.add_points(X, fX)
    o = orc.OracleRBF(d, lb, ub, kernel, "linear")
    o.add_points(X, fX)
    m = 4096 + 37
    q = lb + (ub - lb) * rng.random((m, d))
    q[:64] = X[rng.integers(0, n, 64)]
    for i, eps in enumerate((1e-7, 1e-5, 1e-3)):
        sl = slice(64 + 64 * i, 128 + 64 * i)
        q[sl] = np.clip(X[rng.integers(0, n, 64)] + eps * (ub - lb) * rng.standard_normal((64, d)), lb, ub)
    got = s.predict_deriv(q)
    pick = np.r_[0:320, m - 60:m]
    ref = o.predict_deriv(q[pick])
    err = relmax(got[pick], ref)
    s.handle.set_option(2 + 4, 0)  # B200RBF_OPT_MMA_DIST = 0: the direct-difference tile kernel on the same queries
    direct = s.predict_deriv(q)
    print("%s d=%d n=%d: tensor path vs oracle %.2e, direct tile kernel vs oracle %.2e" % (kernel, d, n, err, relmax(direct[pick], ref)))
    assert err < TOL and relmax(direct[pick], ref) < TOL
    assert relmax(got, direct) < TOL
