"""GPU (B200): seeded random sweep of the whole evaluation path against the oracle.

Every case draws its own dimension, cloud (uniform / clustered / lattice with duplicates), box (isotropic or not),
kernel + tail, call size (1 .. 9000 rows: all three center-split regimes of the DMMA score kernel), pending points,
weights and dtol, then checks predict, predict_deriv, the raw scores and the picks.  The hand-written cases of
test_gpu_parity.py each pin one mechanism; this file is there for the combinations nobody thought of (the tiny-pair
threshold bug of round 2 was found by exactly such a combination: 1e-6-wide clusters at a call size above 7104).
"""
import numpy as np
import pytest

from conftest import relmax

pytestmark = pytest.mark.gpu

KERNELS = {"cubic": "CubicKernel", "tps": "TPSKernel", "linear": "LinearKernel"}
TAILS = {"linear": "LinearTail", "constant": "ConstantTail"}
COMBOS = [("cubic", "linear"), ("tps", "linear"), ("linear", "linear"), ("linear", "constant")]


def draw_cloud(rng, kind, k, lb, ub, centers):
    d = len(lb)
    if kind == "uniform":
        return lb + (ub - lb) * rng.random((k, d))
    if kind == "clustered":
        which = rng.integers(0, len(centers), k)
        rad = 10.0 ** rng.uniform(-6.5, -0.5, (k, 1)) * (ub - lb)
        return np.clip(centers[which] + rad * rng.standard_normal((k, d)) / np.sqrt(d), lb, ub)
    # lattice: few distinct values per coordinate -> many exactly coincident / exactly equidistant points
    return lb + (ub - lb) * rng.integers(0, 4, (k, d)) / 3.0


@pytest.mark.parametrize("seed", range(160))
def test_random_case_against_oracle(seed):
    import pysot_b200 as pb
    from oracle import rbf_oracle as orc

    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice([1, 2, 3, 5, 8, 9, 10, 13, 20, 30, 31, 33, 47, 50, 64, 70]))
    kernel, tail = COMBOS[int(rng.integers(0, len(COMBOS)))]
    kind = ["uniform", "clustered", "lattice"][int(rng.integers(0, 3))]
    n = int(rng.integers(d + 2, 900))
    m = int(rng.choice([1, 2, 15, 16, 17, 100, 1000, 3000, 7200, 9000]))
    if rng.random() < 0.5:
        lb, ub = np.zeros(d), np.ones(d)
    else:
        lb = -10.0 ** rng.uniform(-1, 2, d)
        ub = 10.0 ** rng.uniform(-1, 2, d)
    centers = lb + (ub - lb) * rng.random((6, d))
    X = draw_cloud(rng, kind, n, lb, ub, centers)
    if kind == "lattice":  # the fit needs distinct points: keep the duplicates for the CANDIDATES only
        X = np.unique(X, axis=0)
        if X.shape[0] < d + 2:
            X = lb + (ub - lb) * rng.random((d + 2, d))
        n = X.shape[0]
    fX = np.sin(X.sum(1) / (1.0 + np.abs(ub).max()))[:, None]
    cand = draw_cloud(rng, kind, m, lb, ub, centers)
    k = min(m, 3)
    cand[:k] = X[:k]
    npend = int(rng.integers(0, 5))
    Xpend = draw_cloud(rng, kind, npend, lb, ub, centers) if npend else None
    p = int(rng.integers(1, min(m, 4) + 1))
    weights = [float(w) for w in rng.choice([0.0, 0.3, 0.5, 0.8, 0.95, 1.0], p)]
    dtol = float(rng.choice([1e-3, 1e-6, 1e-1]))

    o = orc.OracleRBF(d, lb, ub, kernel, tail)
    o.add_points(X, fX)
    o.c = rng.standard_normal((n + o.ntail, 1))  # evaluation path only: no (possibly singular) fit
    o.updated = True
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=getattr(pb, KERNELS[kernel])(), tail=getattr(pb, TAILS[tail])(d),
                             exact_dist=bool(rng.random() < 0.15))
    s.add_points(X, fX)
    s.set_coefficients(o.c)
    tag = dict(seed=seed, d=d, kernel=kernel, tail=tail, cloud=kind, n=n, m=m, p=p, pend=npend, dtol=dtol)

    assert relmax(s.predict(cand), o.predict(cand)) < 1e-10, tag
    q = cand[: min(m, 300)]
    if kernel != "linear" or kind == "uniform":  # the linear kernel's gradient is singular at coincident points
        gq = q if kernel != "linear" else q[k:]
        if gq.shape[0]:
            assert relmax(s.predict_deriv(gq), o.predict_deriv(gq)) < 1e-10, tag

    tr = {}
    orc.weighted_distance_merit(p, o, X, fX, cand.copy(), weights, Xpend, dtol, trace=tr)
    idx, _, fv, dm, _ = pb.merit_select_indices(p, s, X, cand, weights, Xpend, dtol, want_scores=True)
    ref_d = tr["dmerit0"].ravel()
    assert relmax(fv, tr["fvals_raw"].ravel()) < 1e-10, tag
    assert np.all(dm[:k] == 0.0), tag
    idx2, _, fv2, dm2, _ = pb.merit_select_indices(p, s, X, cand, weights, Xpend, dtol, want_scores=True)
    assert np.array_equal(idx, idx2) and np.array_equal(fv, fv2, equal_nan=True) and np.array_equal(dm, dm2), tag  # run to run: bits
    # per-candidate accuracy of the distance: relative 1e-9, plus the rounding of the unit-box map itself -- with an
    # isotropic box the library derives the distance from the unit-box one (x dscale), and (x - lb) / (ub - lb) rounds
    # each coordinate to eps |u|, i.e. eps * box width in the original space (the reference's cdist subtracts the
    # original coordinates, which is exact for close points).  Exactly coincident points still give exactly 0.
    pos = ref_d > 0
    atol = 8.0 * np.finfo(float).eps * float(np.max(ub - lb)) * np.sqrt(d)
    assert np.all(np.abs(dm - ref_d) <= 1e-9 * ref_d + atol), (tag, float(np.max(np.abs(dm - ref_d) - 1e-9 * ref_d)), atol)
    assert np.all(dm[~pos] == 0.0), tag
    for i in range(p):  # picks: identical, except where the reference's own merits tie within 1e-12
        if idx[i] == tr["picked"][i]:
            continue
        mi = tr["merit"][i].ravel()
        a, b = mi[idx[i]], mi[tr["picked"][i]]
        assert (np.isinf(a) and np.isinf(b)) or abs(a - b) <= 1e-12, (tag, i, int(idx[i]), int(tr["picked"][i]), a, b)
        break  # after a tie the two selections legitimately diverge


@pytest.mark.parametrize("seed", range(16))
def test_random_growth_sequence_against_fresh_oracle_fit(seed):
    """Points arrive in random batches (1 .. 9 rows) on top of a random base fit, as under a strategy; whichever of fit /
    extension / refit the library chooses, predictions and gradients must match the oracle's FRESH fit of all points."""
    import pysot_b200 as pb
    from oracle import rbf_oracle as orc

    rng = np.random.default_rng(5000 + seed)
    d = int(rng.choice([2, 5, 10, 20, 30]))
    kernel, tail = COMBOS[int(rng.integers(0, 2))] if rng.random() < 0.8 else COMBOS[int(rng.integers(2, 4))]
    lb, ub = -1.0 - rng.random(d), 1.0 + 3.0 * rng.random(d)
    n0 = int(rng.integers(2 * (d + 1), 400))
    total = n0 + int(rng.integers(5, 120))
    capped = rng.random() < 0.25
    X = lb + (ub - lb) * rng.random((total, d))
    fX = np.cos(X.sum(1)) + 0.1 * X[:, 0] ** 2
    kw = dict(output_transformation=pb.median_capping) if capped else {}
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=getattr(pb, KERNELS[kernel])(), tail=getattr(pb, TAILS[tail])(d), **kw)
    s.add_points(X[:n0], fX[:n0])
    s._fit()
    kinds, have = [s.fit_kind], n0
    q = lb + (ub - lb) * rng.random((64, d))
    while have < total:
        k = min(int(rng.integers(1, 10)), total - have)
        s.add_points(X[have:have + k], fX[have:have + k])
        have += k
        if rng.random() < 0.7:
            s.predict(q[:1])  # lazy refit / extension now; otherwise the next batch is appended first
            kinds.append(s.fit_kind)
    o = orc.OracleRBF(d, lb, ub, kernel, tail, output_transformation=orc.median_capping if capped else None)
    o.add_points(X, fX[:, None])
    tag = dict(seed=seed, d=d, kernel=kernel, tail=tail, n0=n0, total=total, capped=capped, kinds=sorted(set(kinds)))
    assert relmax(s.predict(q), o.predict(q)) < 1e-10, tag
    if kernel != "linear":
        assert relmax(s.predict_deriv(q), o.predict_deriv(q)) < 1e-10, tag
    assert relmax(s.predict(X[:50]), o.predict(X[:50])) < 1e-10, tag
