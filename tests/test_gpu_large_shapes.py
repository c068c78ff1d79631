"""GPU parity AT THE BENCHMARKED SHAPES against golden vectors produced by the live reference
(tools/make_large_goldens.py, build container): BASELINE config #3 (TPS, d = 10, N = 5 000 / 20 000: fit + predict +
predict_deriv, both solvers) and config #4 exactly as bench.py builds it (d = 50, n = 20 000, cubic + linear tail,
coefficients from a real fit, default DMMA score kernel, a 2^17-row candidate slice drawn through the reference's
candidate_dycors RNG path).

Bars (north_star): predictions / gradients max|delta| <= 1e-10 max|ref|; minimum distances 1e-12; picks identical
unless the reference's best and second-best merit are within 1e-12.
"""
import hashlib
import os

import numpy as np
import pytest

import pysot_b200 as pb
from pysot_b200 import _lib
from conftest import GOLDEN, Problem, relmax

pytestmark = pytest.mark.gpu


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("%s not generated" % name)
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("n", [5000, 20000])
@pytest.mark.parametrize("method", [0, 1], ids=["auto-cholesky", "lu"])
def test_config3_fit_predict_deriv_vs_reference(n, method):
    g = _load("large_fit_tps10_%d.npz" % n)
    d = int(g["d"])
    rng = np.random.default_rng(int(g["seed"]))
    X = rng.random((n, d))
    fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
    s = pb.GpuRBFInterpolant(dim=d, lb=np.zeros(d), ub=np.ones(d), kernel=pb.TPSKernel(), tail=pb.LinearTail(d))
    s.handle.set_option(_lib.OPT_FIT_METHOD, method)
    s.add_points(X, fX)
    pred = s.predict(g["q"])
    assert s.handle.last_fit_method == ("lu" if method == 1 else "cholesky")
    grad = s.predict_deriv(g["q"][:32])
    ep, eg = relmax(pred, g["pred"]), relmax(grad, g["grad"])
    print("config#3 n=%d %s: predict %.2e, predict_deriv %.2e (fit %.4f s)" % (n, s.handle.last_fit_method, ep, eg, s.fit_seconds))
    assert ep < 1e-10 and eg < 1e-10, (ep, eg)


def _config4():
    g = _load("large_config4.npz")
    d, n, m = int(g["d"]), int(g["n"]), int(g["m"])
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
    prob = Problem(np.zeros(d), np.ones(d))
    np.random.seed(0)
    cand = pb.generate_dycors(prob, X, fX, prob_perturb=0.4, sampling_radius=0.2, num_cand=m)
    assert hashlib.sha256(np.ascontiguousarray(cand).tobytes()).hexdigest() == str(g["cand_sha256"]), \
        "host-generated candidate slice differs from the reference's"
    return g, X, fX, cand


def _check_scores(tag, g, idx, fv, dm):
    ef, ed = relmax(fv, g["fvals"]), relmax(dm, g["dmerit"])
    print("config#4 %s: fvals %.2e, dmerit %.2e, picks %s (reference %s, margins %s)"
          % (tag, ef, ed, idx.tolist(), g["picked"].tolist(), g["margins"].tolist()))
    assert ef < 1e-10 and ed < 1e-12, (tag, ef, ed)
    for i, (a, b) in enumerate(zip(idx.tolist(), g["picked"].tolist())):
        if a != b:
            assert g["margins"][i] <= 1e-12, (tag, i, a, b, float(g["margins"][i]))
            break  # after a legitimate tie the later picks see different distance updates


def test_config4_bench_shape_default_kernel_real_fit():
    """The kernel bench.py times (score_mma_kernel, norm-expansion distances), coefficients from the GPU's own fit."""
    g, X, fX, cand = _config4()
    s = pb.GpuRBFInterpolant(dim=50, lb=np.zeros(50), ub=np.ones(50))
    s.add_points(X, fX)
    idx, merit, fv, dm, st = pb.merit_select_indices(4, s, X, cand, list(g["weights"]), g["Xpend"], 1e-3, want_scores=True)
    assert s.handle.last_fit_method == "cholesky"
    _check_scores("DMMA kernel + GPU fit", g, idx, fv, dm)
    assert relmax(merit, g["merit"]) < 1e-9
    # the fit at the bench shape against the reference's lu_factor solution, through predictions at the centers' scale
    assert relmax(s.predict(cand[:512]), g["fvals"][:512].reshape(-1, 1)) < 1e-10


@pytest.mark.parametrize("variant", ["mma", "direct", "exact", "int8"])
def test_config4_bench_shape_reference_coefficients(variant):
    """Scoring isolated from fitting: the reference's own coefficient vector is loaded into the handle."""
    g, X, fX, cand = _config4()
    s = pb.GpuRBFInterpolant(dim=50, lb=np.zeros(50), ub=np.ones(50), exact_dist=(variant == "exact"))
    if variant == "direct":
        s.handle.set_option(_lib.OPT_MMA_DIST, 0)
    if variant == "int8":  # dot products on the INT8 tensor cores (score_i8.cu)
        s.handle.set_option(_lib.OPT_I8_DIST, 1)
    s.add_points(X, fX)
    s.set_coefficients(g["c"])
    idx, merit, fv, dm, st = pb.merit_select_indices(4, s, X, cand, list(g["weights"]), g["Xpend"], 1e-3, want_scores=True)
    _check_scores(variant + " kernel + reference coefficients", g, idx, fv, dm)
    if variant == "exact":  # separate multiply and add, sequential: scipy's cdist bit for bit
        assert np.array_equal(dm, g["dmerit"])
