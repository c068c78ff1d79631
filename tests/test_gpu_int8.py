"""GPU: the INT8 tensor-core score kernel (csrc/score_i8.cu: tcgen05.mma.kind::i8, accumulators in TMEM, exact fixed-point
dot products) against the oracle, on every path that can reach it: weighted_distance_merit, predict, host and device
candidates, every kernel, dimensions on both sides of each 32-coordinate k-step up to 128, coincident and near-coincident
points, and the fall-back to the FP64 tensor path when a coordinate leaves the unit box."""
import numpy as np
import pytest

import pysot_b200 as pb
from pysot_b200 import _lib
from conftest import relmax

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def orc():
    from oracle import rbf_oracle

    return rbf_oracle


def build(orc, kernel, d, n, seed, lb=None, ub=None):
    rng = np.random.default_rng(seed)
    lb = -np.ones(d) * 2 if lb is None else lb
    ub = np.ones(d) * 3 if ub is None else ub
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(((X - lb) / (ub - lb)).sum(1) * 2)[:, None]
    K = {"cubic": pb.CubicKernel, "tps": pb.TPSKernel, "linear": pb.LinearKernel}[kernel]
    tail = "constant" if kernel == "linear" else "linear"
    T = pb.ConstantTail if kernel == "linear" else pb.LinearTail
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=K(), tail=T(d))
    s.add_points(X, fX)
    o = orc.OracleRBF(d, lb, ub, kernel, tail)
    o.add_points(X, fX)
    return rng, lb, ub, X, fX, s, o


@pytest.mark.parametrize("kernel,d,n", [("cubic", 50, 1500), ("tps", 41, 700), ("linear", 64, 300), ("cubic", 65, 400),
                                        ("tps", 100, 350), ("cubic", 128, 300), ("cubic", 7, 130), ("tps", 32, 260), ("cubic", 33, 300)])
def test_int8_kernel_matches_oracle(orc, kernel, d, n):
    rng, lb, ub, X, fX, s, o = build(orc, kernel, d, n, d)
    m = 6000
    cand = lb + (ub - lb) * rng.random((m, d))
    cand[:40] = X[:40]                                            # coincident with centers: r = 0 exactly
    cand[40:80] = np.clip(X[40:80] + 1e-7 * (ub - lb) * rng.standard_normal((40, d)), lb, ub)  # the near-pair path
    cand[80] = lb
    cand[81] = ub                                                 # the corners of the box: u = 0 and u = 1 exactly
    Xpend = lb + (ub - lb) * rng.random((3, d))
    w = [0.3, 0.5, 0.8, 0.95]
    s.handle.set_option(_lib.OPT_I8_DIST, 1)
    idx, merit, fv, dm, st = pb.merit_select_indices(4, s, X, cand, w, Xpend, want_scores=True)
    tr = {}
    orc.weighted_distance_merit(4, o, X, fX, cand.copy(), w, Xpend, 1e-3, trace=tr)
    ef, ed = relmax(fv, tr["fvals_raw"].ravel()), relmax(dm, tr["dmerit0"].ravel())
    print("int8 %s d=%d n=%d: fvals %.2e dmerit %.2e" % (kernel, d, n, ef, ed))
    assert ef < TOL and ed < 1e-12, (ef, ed)
    assert idx.tolist() == tr["picked"].tolist()
    assert np.all(dm[:40] == 0.0)  # coincident points: distance exactly zero (recomputed directly)
    # predict: same kernel without the distance outputs; host array and device tensor
    assert relmax(s.predict(cand[:3000]), o.predict(cand[:3000])) < TOL
    import torch

    t = torch.from_numpy(cand).cuda()
    assert relmax(s.predict_device(t).cpu().numpy(), o.predict(cand)) < TOL


def test_auto_rule_and_fallback_outside_the_unit_box(orc):
    rng, lb, ub, X, fX, s, o = build(orc, "cubic", 44, 500, 3)
    m = 9000  # auto mode takes the int8 kernel from 8192 rows on
    cand = lb + (ub - lb) * rng.random((m, 44))
    s.predict(cand[:10])  # the fit happens here; few rows: the DMMA kernel
    assert s.handle.last_score_kernel == "dmma"
    s.predict(cand)       # the lazy slicing of the centers
    l0 = s.handle.launches
    p_in = s.predict(cand)  # d = 44 >= 40: the int8 kernel by default
    n_in = s.handle.launches - l0
    assert s.handle.last_score_kernel == "int8"
    assert relmax(p_in, o.predict(cand)) < TOL
    far = cand.copy()
    far[17] = ub + 0.3 * (ub - lb)   # outside the box: u > 1 -> flagged, the call is repeated on the DMMA kernel
    far[99] = lb - 2.0 * (ub - lb)
    l0 = s.handle.launches
    p_out = s.predict(far)
    assert s.handle.launches - l0 > n_in and s.handle.last_score_kernel == "dmma"  # it ran twice
    assert relmax(p_out, o.predict(far)) < TOL
    # and through the acquisition entry point
    w = [0.4, 0.9]
    idx, _, fv, dm, _ = pb.merit_select_indices(2, s, X, far, w, None, want_scores=True)
    tr = {}
    orc.weighted_distance_merit(2, o, X, fX, far.copy(), w, None, 1e-3, trace=tr)
    assert relmax(fv, tr["fvals_raw"].ravel()) < TOL and idx.tolist() == tr["picked"].tolist()
    # the next in-range call is served by the int8 kernel again
    l0 = s.handle.launches
    assert relmax(s.predict(cand), o.predict(cand)) < TOL
    assert s.handle.launches - l0 == n_in and s.handle.last_score_kernel == "int8"
    # centers outside the box (points added beyond the bounds): the model stays on the DMMA kernel
    s.add_points(ub + 0.5, 1.0)
    o.add_points(ub + 0.5, 1.0)
    assert relmax(s.predict(cand), o.predict(cand)) < TOL
    assert s.handle.last_score_kernel == "dmma"
    # exact_dist and MMA_DIST = 0 take precedence
    e = pb.GpuRBFInterpolant(dim=44, lb=lb, ub=ub, exact_dist=True)
    e.add_points(X, fX)
    e.handle.set_option(_lib.OPT_I8_DIST, 1)
    _, _, _, dm_e, _ = pb.merit_select_indices(1, e, X, cand, [0.5], None, want_scores=True)
    import scipy.spatial.distance as ssd

    assert np.array_equal(dm_e, ssd.cdist(cand, X).min(axis=1))
