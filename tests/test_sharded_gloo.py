"""CPU, world_size = 2, gloo: the host logic of the sharded selection loop (pysot_b200/sharded.py).

The product backend (GpuShardBackend) calls the C ABI phase functions on a B200; here a numpy backend with the
same interface (built from the oracle's formulas) stands in, so offsets, the MIN all-reduce of
{fmin,-fmax,dmin,-dmax}, the {merit, index, row} all-gather with lowest-index tie breaking and the device-side
winner decision (empty shards included) are exercised end to end and compared with the unsharded oracle.
"""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.spatial.distance as ssd
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class NumpyShardBackend:
    def __init__(self, oracle_rbf):
        self.o = oracle_rbf
        self.device = torch.device("cpu")

    def new(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def _stats(self):
        if self.cand.shape[0] == 0:
            return torch.full((4,), float("inf"), dtype=torch.float64)
        return torch.tensor([self.fmin, -self.fmax, self.dm.min(), -self.dm.max()], dtype=torch.float64)

    def score(self, cand, Xdist, alias):
        self.cand = cand
        if cand.shape[0] == 0:
            return self._stats()
        self.fv = self.o.predict(cand).ravel()
        self.dm = ssd.cdist(cand, Xdist).min(axis=1)
        self.fmin, self.fmax = self.fv.min(), self.fv.max()
        return self._stats()

    def best(self, w, dtol, stats, offset):
        d = self.cand.shape[1]
        out = torch.zeros(2 + d, dtype=torch.float64)
        if self.cand.shape[0] == 0:
            out[0] = float("inf")
            out[1:2].view(torch.int64)[0] = 0x7FFFFFFFFFFFFFFF
            return out
        fmn, fmx, dmn, dmx = stats[0].item(), -stats[1].item(), stats[2].item(), -stats[3].item()
        with np.errstate(invalid="ignore", divide="ignore"):
            fh = np.where(np.isinf(self.fv), self.fv, np.ones_like(self.fv) if fmx == fmn else (self.fv - fmn) / (fmx - fmn))
            dh = np.ones_like(self.dm) if dmx == dmn else (self.dm - dmn) / (dmx - dmn)
            merit = w * fh + (1.0 - w) * (1.0 - dh)
        merit[self.dm < dtol] = np.inf
        j = int(np.argmin(merit))
        out[0] = merit[j]
        out[1:2].view(torch.int64)[0] = j + offset
        out[2:] = torch.from_numpy(self.cand[j])
        return out

    def apply(self, gathered, world, offset, update, result):
        from pysot_b200.sharded import _winner_record

        rec = _winner_record(torch, gathered.view(world, -1))
        result.copy_(rec)
        j = int(rec[1:2].view(torch.int64)[0]) - offset
        if 0 <= j < self.cand.shape[0]:
            self.fv[j] = np.inf
        if not update:
            return None
        if self.cand.shape[0] > 0:
            self.dm = np.minimum(self.dm, ssd.cdist(self.cand, rec[2:].numpy()[None, :]).ravel())
        return self._stats()

    def finish(self):
        pass


def _case(seed=0):
    from oracle import rbf_oracle as orc

    rng = np.random.default_rng(seed)
    d, n, m = 6, 50, 1001
    lb, ub = -np.ones(d), 2 * np.ones(d)
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(X.sum(1))[:, None]
    Xpend = lb + (ub - lb) * rng.random((2, d))
    cand = lb + (ub - lb) * rng.random((m, d))
    cand[700] = cand[3]  # an exact duplicate across shards: the lower global index must win the tie
    o = orc.OracleRBF(d, lb, ub)
    o.add_points(X, fX)
    return orc, o, X, fX, Xpend, cand


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pysot_b200.sharded import sharded_select

        orc, o, X, fX, Xpend, cand = _case()
        bounds = [0, 640, cand.shape[0]]  # uneven shards
        local = cand[bounds[rank]:bounds[rank + 1]]
        w = [0.3, 0.5, 0.8, 0.95]
        idx, rows = sharded_select(NumpyShardBackend(o), local, np.vstack((X, Xpend)), True, w, 4, 1e-3)
        q.put((rank, idx.tolist(), rows.tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_selection_matches_unsharded_oracle():
    orc, o, X, fX, Xpend, cand = _case()
    tr = {}
    ref_pts = orc.weighted_distance_merit(4, o, X, fX, cand.copy(), [0.3, 0.5, 0.8, 0.95], Xpend, 1e-3, trace=tr)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, idx, rows in res:
        assert idx == tr["picked"].tolist(), (rank, idx, tr["picked"])
        assert np.array_equal(np.array(rows), ref_pts)


def _worker3(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pysot_b200.sharded import sharded_select

        orc, o, X, fX, Xpend, cand = _case(2)
        bounds = [0, 400, 400, cand.shape[0]]  # rank 1 holds NO candidates (it must still take part)
        local = cand[bounds[rank]:bounds[rank + 1]]
        idx, rows = sharded_select(NumpyShardBackend(o), local, np.vstack((X, Xpend)), True, [0.3, 0.9], 2, 1e-3)
        q.put((rank, idx.tolist(), rows.tolist()))
    finally:
        dist.destroy_process_group()


def test_three_ranks_one_empty_shard():
    orc, o, X, fX, Xpend, cand = _case(2)
    tr = {}
    ref_pts = orc.weighted_distance_merit(2, o, X, fX, cand.copy(), [0.3, 0.9], Xpend, 1e-3, trace=tr)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker3, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, idx, rows in res:
        assert idx == tr["picked"].tolist() and np.array_equal(np.array(rows), ref_pts), (rank, idx)


def test_single_process_path_and_tie_rule():
    from pysot_b200.sharded import _better, sharded_select

    orc, o, X, fX, Xpend, cand = _case(1)
    tr = {}
    orc.weighted_distance_merit(3, o, X, fX, cand.copy(), [0.5, 0.8, 0.95], Xpend, 1e-3, trace=tr)
    idx, rows = sharded_select(NumpyShardBackend(o), cand, np.vstack((X, Xpend)), True, [0.5, 0.8, 0.95], 3, 1e-3)
    assert idx.tolist() == tr["picked"].tolist() and np.array_equal(rows, cand[idx])
    nan = float("nan")
    assert _better(1.0, 5, 1.0, 7) and not _better(1.0, 7, 1.0, 5) and _better(0.5, 9, 1.0, 0)
    assert _better(nan, 9, 0.0, 0) and not _better(0.0, 0, nan, 9) and _better(nan, 1, nan, 2)
