"""GPU, needs >= 2 devices on the box (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):

  * one process, several devices: GpuRBFInterpolant(devices=[...]) shards weighted_distance_merit / candidate_dycors over its
    replicas and returns the single-device picks (pysot_b200/multi.py);
  * a handle on a second device works after the first device was used in the same process (the > 48 KB
    shared-memory opt-in is per device);
  * one process per GPU (torchrun, NCCL): tools/check_sharded.py -- N ranks' picks == the single-GPU picks, uneven
    shards, duplicate candidates across shards.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, Problem, relmax

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


needs2 = pytest.mark.skipif("_ngpu() < 2", reason="needs at least two GPUs")


def _case(d=12, n=400, m=60000, seed=0):
    rng = np.random.default_rng(seed)
    lb, ub = -np.ones(d), 2.0 * np.ones(d)
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(X.sum(1))[:, None]
    Xpend = lb + (ub - lb) * rng.random((3, d))
    cand = lb + (ub - lb) * rng.random((m, d))
    cand[m - 11] = cand[5]  # a duplicate in another device's block: the lower global index must win
    return lb, ub, X, fX, Xpend, cand


@needs2
def test_second_device_after_first():
    import pysot_b200 as pb
    from oracle import rbf_oracle as orc

    lb, ub, X, fX, Xpend, cand = _case(d=10, n=700, m=4000)
    o = orc.OracleRBF(10, lb, ub)
    o.add_points(X, fX)
    for dev in (0, 1, 0):
        s = pb.GpuRBFInterpolant(dim=10, lb=lb, ub=ub, device=dev)
        s.add_points(X, fX)
        assert relmax(s.predict(cand[:256]), o.predict(cand[:256])) < 1e-10
        s.handle.set_option(pb._lib.OPT_FIT_METHOD, 1)  # the LU's 109 KB GEMM and panel kernels as well
        s.updated = False
        s._fit_Xu = None
        assert relmax(s.predict(cand[:256]), o.predict(cand[:256])) < 1e-10
        idx, _ = pb.merit_select_indices(2, s, X, cand, [0.3, 0.8], Xpend)
        assert len(idx) == 2


@needs2
def test_device_group_matches_single_device():
    import pysot_b200 as pb

    lb, ub, X, fX, Xpend, cand = _case()
    w = [0.3, 0.5, 0.8, 0.95]
    one = pb.GpuRBFInterpolant(dim=12, lb=lb, ub=ub, device=0)
    one.add_points(X, fX)
    ref_idx, ref_merit = pb.merit_select_indices(4, one, X, cand, w, Xpend)
    for devs in ([0, 1], list(range(_ngpu()))):
        s = pb.GpuRBFInterpolant(dim=12, lb=lb, ub=ub, devices=devs)
        s.add_points(X, fX)
        idx, merit = pb.merit_select_indices(4, s, X, cand, w, Xpend)
        assert idx.tolist() == ref_idx.tolist(), (devs, idx, ref_idx)
        assert relmax(merit, ref_merit) < 1e-12
        pts = pb.weighted_distance_merit(4, s, X, fX, cand, w, Xpend)
        assert np.array_equal(pts, cand[ref_idx])
        # through the reference's entry point, seeded: same proposals as the single-device surrogate
        prob = Problem(lb, ub)
        np.random.seed(7)
        a = pb.candidate_dycors(num_pts=3, opt_prob=prob, surrogate=s, X=X, fX=fX, weights=w, prob_perturb=0.5, num_cand=20000)
        np.random.seed(7)
        b = pb.candidate_dycors(num_pts=3, opt_prob=prob, surrogate=one, X=X, fX=fX, weights=w, prob_perturb=0.5, num_cand=20000)
        assert np.array_equal(a, b)
        s.add_points(cand[:3], np.zeros(3))  # a refit: the replicas must receive the new model
        one.add_points(cand[:3], np.zeros(3))
        X2 = np.vstack((X, cand[:3]))
        i2, _ = pb.merit_select_indices(2, s, X2, cand, w, None)
        r2, _ = pb.merit_select_indices(2, one, X2, cand, w, None)
        assert i2.tolist() == r2.tolist()
        one.reset(); one.add_points(X, fX)
        s.device_group.close()


@needs2
def test_torchrun_nccl_ranks_match_single_gpu():
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
