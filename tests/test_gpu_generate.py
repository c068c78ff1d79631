"""GPU: device-side candidate generation (b200rbf_generate).

The device streams cannot equal numpy's / scipy's, so parity here is DISTRIBUTIONAL: every perturbed coordinate must
follow scipy's truncnorm with the reference's parameters (candidate_dycors.py:80-86), the DYCORS mask must have the
reference's statistics including its quirk (a forced perturbation never lands on the last coordinate, :78), integer
variables are rounded and clipped (utils.py:68-93), and selection on a device-generated set must equal the oracle's
selection on the very same set.
"""
import numpy as np
import pytest
import scipy.stats as st

from conftest import Problem

pytestmark = pytest.mark.gpu


def setup(d=8, n=40, seed=0, int_var=()):
    import pysot_b200 as pb

    rng = np.random.default_rng(seed)
    lb, ub = -2.0 * np.ones(d), 3.0 * np.ones(d)
    ub[1] = 10.0
    prob = Problem(lb, ub, int_var)
    X = lb + (ub - lb) * rng.random((n, d))
    if len(int_var):
        X[:, list(int_var)] = np.round(X[:, list(int_var)])
    fX = ((X - 0.3) ** 2).sum(1, keepdims=True)
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
    s.add_points(X, fX)
    return pb, prob, s, X, fX


def test_srbf_marginals_follow_truncnorm():
    pb, prob, s, X, fX = setup()
    m = 200000
    cand = pb.generate_on_device("srbf", prob, s, X, fX, sampling_radius=0.2, num_cand=m, seed=1).cpu().numpy()
    xbest = X[np.argmin(fX)]
    assert cand.shape == (m, prob.dim) and np.all(cand >= prob.lb) and np.all(cand <= prob.ub)
    for k in range(prob.dim):
        sig = 0.2 * (prob.ub[k] - prob.lb[k])
        dist = st.truncnorm(a=(prob.lb[k] - xbest[k]) / sig, b=(prob.ub[k] - xbest[k]) / sig, loc=xbest[k], scale=sig)
        assert st.kstest(cand[:50000, k], dist.cdf).pvalue > 1e-4, k
    # coordinates are independent of each other and of the row index
    c = np.corrcoef(cand[:, :4].T)
    assert np.max(np.abs(c - np.eye(4))) < 0.02
    # same seed -> same set; another seed -> another set
    again = pb.generate_on_device("srbf", prob, s, X, fX, sampling_radius=0.2, num_cand=m, seed=1).cpu().numpy()
    other = pb.generate_on_device("srbf", prob, s, X, fX, sampling_radius=0.2, num_cand=m, seed=2).cpu().numpy()
    assert np.array_equal(cand, again) and not np.array_equal(cand, other)


def test_dycors_mask_statistics_and_quirk():
    pb, prob, s, X, fX = setup(d=6)
    xbest = X[np.argmin(fX)]
    m = 300000
    for p in (0.5, 0.05):
        cand = pb.generate_on_device("dycors", prob, s, X, fX, prob_perturb=p, num_cand=m, seed=7).cpu().numpy()
        pert = cand != xbest
        k = pert.sum(1)
        assert k.min() >= 1  # no candidate equals xbest (candidate_dycors.py:77-78)
        # P(no coordinate drawn) = (1-p)^d ; those rows get exactly one forced coordinate, never the last one
        p0 = (1 - p) ** prob.dim
        frac = pert.mean(0)
        expect_first = p + p0 / (prob.dim - 1)
        assert np.allclose(frac[:-1], expect_first, atol=4e-3), (p, frac)
        assert abs(frac[-1] - p) < 4e-3, (p, frac)
        single = pert[k == 1]
        # among single-coordinate rows the last coordinate is under-represented exactly as in the reference
        w_last = p * (1 - p) ** (prob.dim - 1)
        w_other = w_last + p0 / (prob.dim - 1)
        assert abs(single[:, -1].mean() - w_last / (w_last + (prob.dim - 1) * w_other)) < 6e-3
    one = pb.generate_on_device("dycors", prob, s, X, fX, prob_perturb=0.01, num_cand=1000, subset=[2], seed=3).cpu().numpy()
    assert np.all(one[:, 2] != xbest[2]) and np.all(np.delete(one, 2, axis=1) == np.delete(xbest, 2))  # |subset| == 1


def test_uniform_subset_and_integer_variables():
    pb, prob, s, X, fX = setup(d=5, int_var=(1, 3))
    xbest = X[np.argmin(fX)]
    cand = pb.generate_on_device("uniform", prob, s, X, fX, subset=[0, 1, 4], num_cand=100000, seed=5).cpu().numpy()
    assert np.all(cand[:, [2, 3]] == xbest[[2, 3]])
    assert st.kstest(cand[:, 0], st.uniform(prob.lb[0], prob.ub[0] - prob.lb[0]).cdf).pvalue > 1e-4
    assert np.all(cand[:, 1] == np.round(cand[:, 1])) and cand[:, 1].min() == prob.lb[1] and cand[:, 1].max() == prob.ub[1]
    cnt = np.bincount((cand[:, 1] - prob.lb[1]).astype(int))
    assert abs(cnt[1] / cnt[0] - 2.0) < 0.1  # end bins collect half a unit, inner bins a full unit (round + clip)
    srbf = pb.generate_on_device("srbf", prob, s, X, fX, sampling_radius=0.01, num_cand=50000, seed=6).cpu().numpy()
    assert np.all(srbf[:, [1, 3]] == np.round(srbf[:, [1, 3]]))
    assert srbf[:, 1].std() > 0.5  # sigma is at least 1 for integer variables (candidate_srbf.py:100-103)


def test_selection_on_device_generated_candidates_matches_oracle():
    from oracle import rbf_oracle as orc

    pb, prob, s, X, fX = setup(d=10, n=80, seed=3)
    Xpend = prob.lb + (prob.ub - prob.lb) * np.random.default_rng(1).random((3, prob.dim))
    w = [0.3, 0.5, 0.8, 0.95]
    cand_t = pb.generate_on_device("dycors", prob, s, X, fX, prob_perturb=0.4, num_cand=20000, seed=11)
    pts = pb.weighted_distance_merit(4, s, X, fX, cand_t, w, Xpend)
    o = orc.OracleRBF(prob.dim, prob.lb, prob.ub)
    o.add_points(X, fX)
    ref = orc.weighted_distance_merit(4, o, X, fX, cand_t.cpu().numpy(), w, Xpend)
    assert np.array_equal(pts, ref)
    # the switch used by the strategies: candidate_* draw on the device, nothing is generated on the host
    pb.set_candidate_generator("device", seed=123)
    try:
        state = np.random.get_state()[1].copy()
        a = pb.candidate_dycors(num_pts=2, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=w[:2], prob_perturb=0.3,
                                Xpend=Xpend, num_cand=5000)
        b = pb.candidate_srbf(num_pts=2, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=w[:2], Xpend=Xpend, num_cand=5000)
        c = pb.candidate_uniform(num_pts=2, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=w[:2], Xpend=Xpend,
                                 num_cand=5000)
        assert np.array_equal(state, np.random.get_state()[1])  # the host RNG was not consumed
        for pts in (a, b, c):
            assert pts.shape == (2, prob.dim) and np.all(pts >= prob.lb) and np.all(pts <= prob.ub)
    finally:
        pb.set_candidate_generator("host")


def test_sixteen_million_device_generated_candidates():
    """BASELINE config #4's upper range on one GPU without the host: 2^24 candidates (6.7 GB) are drawn, scored against
    2 000 centers and selected on the device; size-independent properties are checked with torch on the same tensors."""
    import torch

    import pysot_b200 as pb

    rng = np.random.default_rng(0)
    d, n, m = 50, 2000, 1 << 24
    prob = Problem(np.zeros(d), np.ones(d))
    X = rng.random((n, d))
    fX = (np.sin(3 * X.sum(1)) + X[:, 0] ** 2)[:, None]
    s = pb.GpuRBFInterpolant(dim=d, lb=prob.lb, ub=prob.ub)
    s.add_points(X, fX)
    cand = pb.generate_on_device("dycors", prob, s, X, fX, prob_perturb=0.4, num_cand=m, seed=5)
    assert cand.shape == (m, d) and float(cand.min()) >= 0.0 and float(cand.max()) <= 1.0
    idx, merit, fv, dm, st = pb.merit_select_indices(4, s, X, cand, [0.3, 0.5, 0.8, 0.95], None, want_scores=True)
    assert len(set(idx.tolist())) == 4 and idx.min() >= 0 and idx.max() < m
    assert st[0] == fv.min() and st[1] == fv.max() and st[2] == dm.min() and st[3] == dm.max()
    # spot-check rows from the far end of the index range against a torch fp64 evaluation of the same formula
    rows = torch.tensor([0, 1, m // 2, m - 2, m - 1] + idx.tolist(), device=cand.device)
    Xt = torch.from_numpy(X).to(cand.device)
    r = torch.cdist(cand[rows], Xt)
    c = torch.from_numpy(s.c).to(cand.device)
    ref_f = (r ** 3) @ c[d + 1:] + c[0] + cand[rows] @ c[1:d + 1]
    ref_d = r.min(dim=1).values
    got_f = torch.from_numpy(fv)[rows.cpu()]
    got_d = torch.from_numpy(dm)[rows.cpu()]
    assert float((ref_f.cpu().ravel() - got_f).abs().max()) < 1e-9 * float(ref_f.abs().max())
    assert float((ref_d.cpu() - got_d).abs().max()) < 1e-12
    # every pick is at least dtol away from every evaluated point and from the earlier picks
    P = cand[torch.as_tensor(idx, device=cand.device)]
    assert float(torch.cdist(P, Xt).min()) >= 1e-3 and float(torch.pdist(P).min()) >= 1e-3
