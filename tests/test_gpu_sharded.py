"""GPU: the phase API used by the multi-GPU path (b200rbf_shard_*), driven through sharded_select at world size 1,
must reproduce the single-call select; with torchrun it is exercised by bench.py --gpus N."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(m=30000, d=12, n=300, seed=0):
    import pysot_b200 as pb

    rng = np.random.default_rng(seed)
    lb, ub = np.zeros(d), np.ones(d)
    X = rng.random((n, d))
    fX = np.sin(3 * X.sum(1))[:, None]
    Xpend = rng.random((3, d))
    xbest = X[np.argmin(fX)]
    cand = np.where(rng.random((m, d)) < 0.4, np.clip(xbest + 0.2 * rng.standard_normal((m, d)), 0, 1), xbest)
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub)
    s.add_points(X, fX)
    return pb, s, X, fX, Xpend, cand


@pytest.mark.parametrize("device_input", [False, True])
def test_phase_api_matches_single_call(device_input):
    import torch
    from pysot_b200.sharded import GpuShardBackend, sharded_select, sharded_weighted_distance_merit

    pb, s, X, fX, Xpend, cand = _setup()
    w = [0.3, 0.5, 0.8, 0.95]
    ref_idx, ref_merit = pb.merit_select_indices(4, s, X, cand, w, Xpend)
    assert len(set(ref_idx.tolist())) == 4
    inp = torch.from_numpy(cand).cuda() if device_input else cand
    be = GpuShardBackend(s)
    idx, rows = sharded_select(be, inp, np.vstack((X, Xpend)), True, w, 4, 1e-3)
    assert idx.tolist() == ref_idx.tolist(), (idx, ref_idx)
    assert np.array_equal(rows, cand[ref_idx])
    pts, idx2 = sharded_weighted_distance_merit(4, s, X, fX, inp, w, Xpend)
    assert idx2.tolist() == ref_idx.tolist() and np.array_equal(pts, cand[ref_idx])


def test_phase_stats_track_the_updates():
    import torch

    pb, s, X, fX, Xpend, cand = _setup(m=5000)
    h = s.handle
    s._fit()
    fv, dm, st = h.score(cand, s.lb, s.ub, np.vstack((X, Xpend)), True, want_scores=True)
    sd = torch.zeros(4, dtype=torch.float64, device="cuda")
    h.shard_stats(sd)
    assert np.array_equal(sd.cpu().numpy(), [st[0], -st[1], st[2], -st[3]])
    pair = torch.zeros(2, dtype=torch.float64, device="cuda")
    h.shard_argmin(0.3, 1e-3, sd, 1000, pair)
    j = int(pair.cpu().view(torch.int64)[1]) - 1000
    row = torch.zeros(s.dim, dtype=torch.float64, device="cuda")
    h.shard_take(j, row)
    assert np.array_equal(row.cpu().numpy(), cand[j])
    h.shard_update(row)
    h.shard_stats(sd)
    dm2 = np.minimum(dm, np.sqrt(((cand - cand[j]) ** 2).sum(1)))
    got = sd.cpu().numpy()
    assert got[2] == 0.0 and abs(-got[3] - dm2.max()) < 1e-12 and got[0] == st[0] and got[1] == -st[1]
