"""Several GPUs behind the drop-in API, ONE process: ``GpuRBFInterpolant(devices=[0, 1, ...])``.

A strategy calls ``candidate_dycors(...)`` / ``weighted_distance_merit(...)`` from the single POAP controller thread
(surrogate_strategy.py:249-299); it cannot be launched under ``torchrun``.  With ``devices=[...]`` the surrogate keeps
one library handle per device (replicas of centers + coefficients, ``b200rbf_load``; the fit itself runs on the first
device: it does not shard), splits the candidate rows into contiguous blocks, scores the blocks concurrently -- one host
thread per device, the ctypes calls release the GIL -- and runs the same selection phases as ``pysot_b200/sharded.py``
with this process playing the part of the collectives: the few doubles per pick ({fmin, -fmax, dmin, -dmax} and the
ranks' {merit, index, row} records) travel through pinned host memory.  Global index = block offset + local index, so
np.argmin's first-minimum rule (candidate_srbf.py:49) holds across devices.

The multi-process path (one rank per GPU, NCCL) is ``pysot_b200/sharded.py``; it is what ``bench.py --gpus N`` times.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib


class DeviceGroup:
    """Replica handles on ``devices`` + the sharded weighted-distance-merit loop over them."""

    def __init__(self, surrogate, devices):
        import torch

        self.torch = torch
        self.s = surrogate
        self.devices = [int(d) for d in devices]
        assert len(self.devices) >= 1 and len(set(self.devices)) == len(self.devices)
        self.handles = {}  # device -> Handle (the surrogate's own handle serves its first device)
        self._model_of = {}  # device -> id of the coefficient array it holds
        self.pool = ThreadPoolExecutor(len(self.devices))

    def handle(self, dev):
        if dev == self.s._device:
            return self.s.handle
        if dev not in self.handles:
            h = _lib.Handle(dev)
            if self.s._exact:
                h.set_option(_lib.OPT_EXACT_DIST, 1)
            self.handles[dev] = h
        return self.handles[dev]

    def sync_model(self):
        """Replicas receive centers + coefficients of the fitted model (``b200rbf_load``), once per fit."""
        self.s._fit()
        key = id(self.s.c)
        n = self.s.num_pts
        for dev in self.devices:
            if dev != self.s._device and self._model_of.get(dev) != key:
                self.handle(dev).load_model(self.s._X[:n], self.s.c, self.s._kid, self.s._tid)
                self._model_of[dev] = key

    def select(self, num_pts, X, cand, weights, Xpend, dtol):
        """Indices (into ``cand``) and merit values of the ``num_pts`` picks; ``cand`` is a host array."""
        t = self.torch
        self.sync_model()
        cand = np.ascontiguousarray(cand, dtype=np.float64)
        m, d = cand.shape
        G = len(self.devices)
        bounds = [(m * g) // G for g in range(G + 1)]
        Xdist = np.vstack((X, Xpend))
        alias = X.shape == self.s.X.shape and np.array_equal(X, self.s.X)
        live = [g for g in range(G) if bounds[g + 1] > bounds[g]]
        w = _lib.f64(list(weights)[:num_pts])
        if w.shape[0] < num_pts:
            raise IndexError("list index out of range")  # weights[i] in candidate_srbf.py:45

        def score(g):
            h = self.handle(self.devices[g])
            _, _, st = h.score(cand[bounds[g]:bounds[g + 1]], self.s.lb, self.s.ub, Xdist, alias)
            return st  # {fmin, fmax, dmin, dmax}

        stats = list(self.pool.map(score, live))
        glob = np.array([min(s[0] for s in stats), -max(s[1] for s in stats), min(s[2] for s in stats),
                         -max(s[3] for s in stats)])
        dev_buf = {}
        for g in live:
            dv = t.device("cuda", self.devices[g])
            dev_buf[g] = dict(stats=t.empty(4, dtype=t.float64, device=dv), rec=t.empty(2 + d, dtype=t.float64, device=dv),
                              gathered=t.empty(len(live) * (2 + d), dtype=t.float64, device=dv),
                              result=t.empty(2 + d, dtype=t.float64, device=dv), new=t.empty(4, dtype=t.float64, device=dv))
        picked, merit = np.empty(num_pts, dtype=np.int64), np.empty(num_pts)
        recs = t.empty((len(live), 2 + d), dtype=t.float64, pin_memory=True)
        for i in range(num_pts):
            gt = t.from_numpy(glob)
            for k, g in enumerate(live):  # local argmin with the global statistics
                b = dev_buf[g]
                b["stats"].copy_(gt, non_blocking=False)
                h = self.handle(self.devices[g])
                h.set_stream(t.cuda.current_stream(b["rec"].device).cuda_stream)
                h.shard_best(float(w[i]), dtol, b["stats"], bounds[g], b["rec"])
                recs[k].copy_(b["rec"], non_blocking=True)
            for g in live:
                t.cuda.synchronize(self.devices[g])
            last = i + 1 == num_pts
            for g in live:  # every device decides on the same gathered records and updates its block
                b = dev_buf[g]
                b["gathered"].copy_(recs.view(-1), non_blocking=True)
                self.handle(self.devices[g]).shard_apply(b["gathered"], len(live), bounds[g], not last, b["result"], b["new"])
            b0 = dev_buf[live[0]]
            res = b0["result"].cpu()
            merit[i] = float(res[0])
            picked[i] = int(res[1:2].view(t.int64)[0])
            if not last:
                new = np.stack([dev_buf[g]["new"].cpu().numpy() for g in live])
                glob[2:] = new[:, 2:].min(axis=0)  # fvals' min / max are computed once (candidate_srbf.py:40)
        for g in live:
            self.handle(self.devices[g]).set_stream(None)
        return picked, merit

    def close(self):
        self.pool.shutdown(wait=False)
        for h in self.handles.values():
            h.close()
        self.handles.clear()
