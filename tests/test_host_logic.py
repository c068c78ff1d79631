"""CPU: host-side mirror of the reference interface (pysot_b200/interface.py, rbf.py, candidates.py)."""
import pickle

import numpy as np
import pytest

from conftest import Problem
import pysot_b200 as pb
from pysot_b200 import candidates as cnd


# ---- kernels / tails: the reference's known answers (tests/test_surrogates.py:37-112)
def test_cubic_kernel():
    k = pb.CubicKernel()
    assert isinstance(k, pb.Kernel) and k.order == 2
    np.testing.assert_allclose(k.eval(np.array(2)), 8)
    np.testing.assert_allclose(k.deriv(np.array(2)), 12)
    X = np.random.rand(10, 3)
    np.testing.assert_allclose(k.eval(X), X ** 3)
    np.testing.assert_allclose(k.deriv(X), 3 * X ** 2)


def test_tps_kernel():
    k = pb.TPSKernel()
    assert isinstance(k, pb.Kernel) and k.order == 2
    np.testing.assert_allclose(k.eval(np.array([2.0])), 4 * np.log(2))
    np.testing.assert_allclose(k.deriv(np.array([2.0])), 2 * (1 + 2 * np.log(2)))
    X = np.random.rand(10, 3)
    np.testing.assert_allclose(k.eval(X.copy()), X ** 2 * np.log(X))
    z = np.zeros(3)
    assert np.all(np.isfinite(k.eval(z))) and z[0] == np.finfo(float).eps  # clamped in place (tps_kernel.py:28)


def test_linear_kernel():
    k = pb.LinearKernel()
    assert isinstance(k, pb.Kernel) and k.order == 1
    X = np.random.rand(10, 3)
    np.testing.assert_allclose(k.eval(X), X)
    np.testing.assert_allclose(k.deriv(X), np.ones((10, 3)))


def test_tails():
    lt = pb.LinearTail(1)
    assert isinstance(lt, pb.Tail) and lt.degree == 1 and lt.dim_tail == 2
    np.testing.assert_allclose(lt.eval(np.array([[2.0]])), np.array([[1, 2]]))
    np.testing.assert_allclose(lt.deriv(np.array([[2.0]])), np.array([[0, 1]]))
    lt = pb.LinearTail(3)
    X = np.random.rand(10, 3)
    np.testing.assert_allclose(lt.eval(X), np.hstack((np.ones((10, 1)), X)))
    np.testing.assert_allclose(lt.deriv(X[0]), np.hstack((np.zeros((3, 1)), np.eye(3))))
    with pytest.raises(ValueError):
        lt.eval(np.random.rand(4, 2))
    ct = pb.ConstantTail(3)
    assert ct.degree == 0 and ct.dim_tail == 1
    np.testing.assert_allclose(ct.eval(X), np.ones((10, 1)))
    np.testing.assert_allclose(ct.deriv(X[0]), np.zeros((3, 1)))
    with pytest.raises(ValueError):
        ct.deriv(np.random.rand(1, 2))


def test_median_capping():
    x = np.array([[1.0], [5.0], [2.0], [9.0], [3.0]])
    y = pb.median_capping(x)
    assert np.array_equal(y.ravel(), [1, 3, 2, 3, 3]) and x[1, 0] == 5.0
    assert pb.identity(x) is x


# ---- surrogate bookkeeping (surrogate.py:33-74, rbf.py:67-95)
def test_constructor_defaults_and_mismatch():
    s = pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3))
    assert isinstance(s, pb.Surrogate)
    assert type(s.kernel).__name__.endswith("CubicKernel") and s.ntail == 4 and s.eta == 1e-6
    # either one missing -> BOTH replaced by the defaults (rbf.py:70-72)
    s = pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3), kernel=pb.TPSKernel())
    assert type(s.kernel).__name__.endswith("CubicKernel")
    with pytest.raises(ValueError, match="Kernel and tail mismatch"):
        pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3), kernel=pb.CubicKernel(), tail=pb.ConstantTail(3))
    pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3), kernel=pb.LinearKernel(), tail=pb.ConstantTail(3))
    with pytest.raises(AssertionError):
        pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3), kernel=pb.CubicKernel(), tail=pb.LinearTail(2))


def test_unsupported_kernel_has_no_fallback():
    class Gauss(pb.Kernel):
        order = 1

        def eval(self, d):
            return np.exp(-d ** 2)

        def deriv(self, d):
            return -2 * d * np.exp(-d ** 2)

    with pytest.raises(NotImplementedError):
        pb.GpuRBFInterpolant(dim=2, lb=np.zeros(2), ub=np.ones(2), kernel=Gauss(), tail=pb.LinearTail(2))


def test_add_points_bookkeeping_and_reset():
    lb, ub = np.array([0.0, -10.0]), np.array([100.0, 10.0])
    s = pb.GpuRBFInterpolant(dim=2, lb=lb, ub=ub)
    s.add_points(np.array([50.0, 0.0]), 3.0)            # float
    s.add_points(np.array([[0.0, -10.0]]), np.array(4.0))  # 0-d
    s.add_points(np.array([[100.0, 10.0], [25.0, 5.0]]), np.array([5.0, 6.0]))  # 1-d
    s.add_points(np.array([[1.0, 1.0]]), np.array([[7.0]]))  # (k, 1)
    assert s.num_pts == 5 and s.X.shape == (5, 2) and s.fX.shape == (5, 1) and not s.updated
    np.testing.assert_allclose(s._X[:3], [[0.5, 0.5], [0, 0], [1, 1]])
    with pytest.raises(AssertionError):
        s.add_points(np.zeros((2, 2)), np.zeros(3))
    s.reset()
    assert s.num_pts == 0 and s.X.shape == (0, 2) and s.fX.shape == (0, 1) and s.c is None and not s.updated


def test_pickle_drops_device_state():
    s = pb.GpuRBFInterpolant(dim=2, lb=np.zeros(2), ub=np.ones(2), kernel=pb.TPSKernel(), tail=pb.LinearTail(2))
    s.add_points(np.random.rand(6, 2), np.random.rand(6))
    s.c, s.updated, s._handle = np.ones((9, 1)), True, object()
    for mod in (pickle, __import__("dill")):
        t = mod.loads(mod.dumps(s))
        assert t._handle is None and t.c is None and t.updated is False
        assert np.array_equal(t.X, s.X) and np.array_equal(t.fX, s.fX) and t._kid == 1 and t._tid == 0


def test_merit_needs_gpu_surrogate():
    class Fake:
        pass

    with pytest.raises(TypeError, match="no CPU fallback"):
        pb.weighted_distance_merit(1, Fake(), np.zeros((2, 2)), np.zeros((2, 1)), np.zeros((4, 2)), [0.5])


# ---- candidate generation consumes the global RNG exactly like the reference (fixtures from the live reference)
def test_generators_match_reference_fixtures(candidate_cases):
    for name in candidate_cases.names:
        g = candidate_cases.get(name)
        prob = Problem(g["lb"], g["ub"], g["int_var"])
        kw = {k[3:]: (v.item() if v.ndim == 0 else v) for k, v in g.items() if k.startswith("kw_")}
        if "subset" in kw:
            kw["subset"] = list(np.atleast_1d(kw["subset"]))
        gen = dict(dycors=cnd.generate_dycors, srbf=cnd.generate_srbf, uniform=cnd.generate_uniform)[str(g["kind"])]
        np.random.seed(int(g["seed"]))
        cand = gen(prob, g["X"], g["fX"], **kw)
        assert np.array_equal(cand, g["cand"]), name


def test_dycors_never_forces_last_coordinate():
    """candidate_dycors.py:78: randint's upper bound is exclusive, so a forced perturbation never lands on the
    last coordinate -- kept for seeded parity."""
    prob = Problem(np.zeros(4), np.ones(4))
    X = np.full((3, 4), 0.5)
    np.random.seed(3)
    cand = cnd.generate_dycors(prob, X, np.array([[1.0], [0.0], [2.0]]), prob_perturb=1e-9, num_cand=2000)
    assert np.all((cand != 0.5).sum(axis=1) == 1)
    assert np.all(cand[:, 3] == 0.5)


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/pySOT"), reason="reference tree not present")
def test_install_rebinds_reference_names():
    from tools import ref_import

    ref_import.load()
    import pySOT.strategy.dycors_strategy as ds
    import pySOT.strategy.srbf_strategy as ss
    import pySOT.strategy.sop_strategy as so
    import sys

    cd = sys.modules["pySOT.auxiliary_problems.candidate_dycors"]  # the package attribute is the function

    orig = ds.candidate_dycors
    done = cnd.install()
    try:
        assert ds.candidate_dycors is cnd.candidate_dycors and so.candidate_dycors is cnd.candidate_dycors
        assert ss.candidate_srbf is cnd.candidate_srbf and cd.weighted_distance_merit is cnd.weighted_distance_merit
        assert so.nd_sorting is cnd.nd_sorting  # SOP's host bottleneck, numpy version with identical output
        assert len(done) >= 10
        # after install() the reference's other surrogates still work -- a non-GPU surrogate is handed to
        # the reference function that install() replaced (here: the stock CPU RBFInterpolant through candidate_dycors)
        from pySOT.surrogate import RBFInterpolant

        prob = Problem(np.zeros(3), np.ones(3))
        rng = np.random.default_rng(0)
        X, fX = rng.random((12, 3)), rng.random((12, 1))
        cpu = RBFInterpolant(dim=3, lb=prob.lb, ub=prob.ub)
        cpu.add_points(X, fX)
        np.random.seed(5)
        want = orig(num_pts=2, opt_prob=prob, surrogate=cpu, X=X, fX=fX, weights=[0.3, 0.8], prob_perturb=0.5, num_cand=50)
        np.random.seed(5)
        got = ds.candidate_dycors(num_pts=2, opt_prob=prob, surrogate=cpu, X=X, fX=fX, weights=[0.3, 0.8],
                                  prob_perturb=0.5, num_cand=50)
        assert np.array_equal(want, got)
    finally:
        cnd.uninstall()
    assert ds.candidate_dycors is orig


def test_candidate_generator_switch():
    assert cnd._GENERATOR["mode"] == "host"
    with pytest.raises(ValueError):
        pb.set_candidate_generator("gpu")
    pb.set_candidate_generator("device", seed=5)
    try:
        assert cnd._GENERATOR == {"mode": "device", "seed": 5, "calls": 0}
    finally:
        pb.set_candidate_generator("host")
    assert cnd._GENERATOR["mode"] == "host"


def test_batched_truncnorm_is_bit_identical_to_scipy_loop():
    """The host generators draw all coordinates in one pass (candidates._truncnorm_rvs_batch); candidate sets and the
    state of numpy's global stream must equal the reference's per-coordinate truncnorm.rvs loop bit for bit."""
    from tools import check_host_generators as chk

    assert cnd._fast_truncnorm_ok()
    assert chk.run(cases=80, seed=5) == 0


# ---- SOP's non-dominated sorting on numpy (pysot_b200/pareto.py)
def _nd_sorting_by_definition(F, nmax):
    """Front peeling straight from the definition (strict domination in every objective, utils.py:389, 465-483)."""
    M, N = F.shape
    ranks = np.full(N, np.inf)
    rest, count, i = list(range(N)), 0, 1
    while count < nmax and rest:
        front = [a for a in rest if not any(np.all(F[:, b] < F[:, a]) for b in rest)]
        for a in front:
            ranks[a] = i
        count += len(front)
        rest = [a for a in rest if a not in front]
        i += 1
    return ranks


def test_nd_sorting_matches_definition_and_reference():
    from pysot_b200.pareto import nd_sorting

    ref = None
    if __import__("os").path.isdir("/root/reference/pySOT"):
        from tools import ref_import

        ref_import.load()
        from pySOT.utils import nd_sorting as ref
    rng = np.random.default_rng(3)
    for it in range(120):
        M, N, nmax = int(rng.choice([1, 2, 2, 3])), int(rng.integers(1, 60)), int(rng.choice([1, 7, 100, 1000]))
        F = rng.random((M, N))
        if it % 3 == 0:
            F = np.round(F * 4) / 4  # ties
        if it % 5 == 0:
            F[:, rng.integers(0, N)] = F[:, rng.integers(0, N)]  # duplicate point
        got = nd_sorting(F.copy(), nmax)
        assert got.dtype == np.float64 and np.array_equal(got, _nd_sorting_by_definition(F, nmax)), it
        if ref is not None:
            assert np.array_equal(got, ref(F.copy(), nmax)), it
    assert nd_sorting(np.zeros((2, 0)), 10).shape == (0,)


def test_fit_decides_between_extension_and_refit():
    """rbf.py:97-161 on the host side: extend when the points the device factorisation was built from are an unchanged
    prefix, refit otherwise or when the library declines (Handle.extend -> None)."""
    calls = []

    class StubHandle:
        last_fit_method = "stub"
        decline = False

        def fit(self, Xu, rhs, kid, tid, eta):
            calls.append(("fit", Xu.shape[0]))
            return np.zeros(Xu.shape[0] + Xu.shape[1] + 1), 0.0

        def extend(self, Xu, rhs, tid, rhs_changed):
            calls.append(("extend", Xu.shape[0], bool(rhs_changed)))
            return None if self.decline else (np.zeros(Xu.shape[0] + Xu.shape[1] + 1), 0.0)

    rng = np.random.default_rng(0)
    s = pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3), output_transformation=pb.median_capping)
    s._handle = StubHandle()
    X, f = rng.random((40, 3)), rng.random((40, 1))
    s.add_points(X[:20], f[:20])
    s._fit()
    s.add_points(X[20:25], f[20:25])
    s._fit()                                  # prefix unchanged -> extend; the median moved -> old values changed
    s._fit()                                  # already up to date: nothing
    s.add_points(X[25], f[25])
    s._handle.decline = True
    s._fit()                                  # the library declines -> full refit
    s._handle.decline = False
    s.reset()
    s.add_points(X[:30], f[:30])
    s._fit()                                  # reset: nothing to extend
    s.add_points(X[30], f[30])
    s._X[3, 0] += 1e-3                        # a resident point moved: refit
    s._fit()
    t = pickle.loads(pickle.dumps(s))
    assert t._fit_Xu is None and t._handle is None
    assert calls == [("fit", 20), ("extend", 25, True), ("extend", 26, True), ("fit", 26), ("fit", 30), ("fit", 31)], calls
    plain = pb.GpuRBFInterpolant(dim=3, lb=np.zeros(3), ub=np.ones(3), incremental=False)
    plain._handle = StubHandle()
    plain.add_points(X[:20], f[:20]); plain._fit(); plain.add_points(X[20], f[20]); plain._fit()
    assert calls[-2:] == [("fit", 20), ("fit", 21)]


def test_add_points_grows_in_place_with_the_reference_bookkeeping():
    """GpuRBFInterpolant.add_points keeps X / _X / fX as prefixes of geometrically growing buffers and maps only the new
    rows; the arrays must equal what surrogate.py:50-74 builds (vstack + re-mapping every row), through scalar / 1-d /
    2-d inputs, growth, pickling, reset() and an outside assignment to X."""
    import pickle

    import pysot_b200.interface as itf
    from pysot_b200.rbf import GpuRBFInterpolant

    class Ref(itf._Surrogate):
        def predict(self, xx):
            raise NotImplementedError

        def predict_deriv(self, xx):
            raise NotImplementedError

    d = 5
    lb, ub = -np.arange(1.0, d + 1), np.arange(2.0, d + 2)
    ref, gpu = Ref(d, lb, ub), GpuRBFInterpolant(d, lb, ub)
    rng = np.random.default_rng(0)

    def same(a, b):
        return (a.num_pts == b.num_pts and np.array_equal(a.X, b.X) and np.array_equal(a._X, b._X)
                and np.array_equal(a.fX, b.fX) and a.fX.shape == b.fX.shape and not b.updated)

    for k in (1, 3, 1, 1, 70, 1, 200, 1):
        x, f = lb + (ub - lb) * rng.random((k, d)), rng.random(k)
        args = (x[0], float(f[0])) if k == 1 else (x, f[:, None] if k == 3 else f)
        ref.add_points(*args)
        gpu.add_points(*args)
        assert same(ref, gpu), k
    back = pickle.loads(pickle.dumps(gpu))
    for s in (ref, back):
        s.add_points(x[0], 1.0)
    assert same(ref, back)
    gpu.reset()
    ref.reset()
    for s in (ref, gpu):
        s.add_points(x, f)
    assert same(ref, gpu)
    gpu.X = gpu.X.copy()  # someone replaced the array: the buffers are re-seated from what is there
    for s in (ref, gpu):
        s.add_points(x, f)
    assert same(ref, gpu)
