"""Candidate scoring sharded over several GPUs: one process per GPU, torch.distributed (NCCL) for the plumbing.

weighted_distance_merit (candidate_srbf.py:8-57) couples candidates only through two min/max rescales and an
argmin, so each rank scores a contiguous block of candidate rows against its replica of the centers and
coefficients, and the ranks exchange only

  * ``{fmin, -fmax, dmin, -dmax}``   one all-reduce(MIN) after scoring, and ``{dmin, -dmax}`` again per extra pick
  * ``{merit, global index, row}``    one all-gather per pick of every rank's local best and the candidate row
                                      behind it (2 + d doubles); each rank then finds the winner on its own device:
                                      lowest merit, lowest index on ties (np.argmin's first-minimum rule,
                                      candidate_srbf.py:49) -- no broadcast, no host round trip between the picks

No candidate data ever crosses NVLink.  The driver below is written against a small backend interface so the
host logic is testable on CPU with gloo; ``GpuShardBackend`` is the product backend (C ABI phase calls).
"""
import numpy as np


class GpuShardBackend:
    """Phase calls of include/b200rbf.h on this rank's handle; all buffers are CUDA tensors, nothing is read back
    between the picks."""

    def __init__(self, surrogate):
        import torch

        self.torch = torch
        self.s = surrogate
        self.h = surrogate.handle
        self.device = torch.device("cuda", surrogate._device)
        self.m = 0

    def new(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype, device=self.device)

    def score(self, cand, Xdist, alias):
        """Fused scoring of the local rows; returns this rank's {fmin, -fmax, dmin, -dmax} (all +inf for an empty shard:
        the neutral element of the MIN all-reduce)."""
        t = self.torch
        self.m = int(cand.shape[0])
        if self.m == 0:
            return t.full((4,), float("inf"), dtype=t.float64, device=self.device)
        self.h.set_stream(t.cuda.current_stream(self.device).cuda_stream)
        if isinstance(cand, np.ndarray):
            self.h.score(cand, self.s.lb, self.s.ub, Xdist, alias)
        else:  # CUDA tensor hand-off
            assert cand.is_cuda and cand.dtype == t.float64 and cand.is_contiguous()
            self.h.score(cand, self.s.lb, self.s.ub, Xdist, alias, m=self.m)
        stats = self.new(4, t.float64)
        self.h.shard_stats(stats)
        return stats

    def best(self, w, dtol, stats, offset):
        """{merit, global index (int64 bits), row[d]} of the local argmin; an empty shard never wins."""
        t = self.torch
        rec = self.new(2 + self.s.dim, t.float64)
        if self.m == 0:
            rec[0] = float("inf")
            rec[1:2].view(t.int64)[0] = _INT64_MAX
            return rec
        self.h.shard_best(w, dtol, stats, offset, rec)
        return rec

    def apply(self, gathered, world, offset, update, result):
        """Winner of the gathered records -> result; mark it taken / fold it into dmerit locally; returns the new local
        {fmin, -fmax, dmin, -dmax} when ``update``."""
        t = self.torch
        if self.m == 0:
            result.copy_(_winner_record(t, gathered.view(world, -1)))
            return t.full((4,), float("inf"), dtype=t.float64, device=self.device) if update else None
        stats = self.new(4, t.float64) if update else None
        self.h.shard_apply(gathered, world, offset, update, result, stats)
        return stats

    def finish(self):
        if self.m > 0:
            self.h.set_stream(None)


_INT64_MAX = 0x7FFFFFFFFFFFFFFF


def _winner_record(torch, rec):
    """np.argmin order over records [merit, index bits, ...]: a NaN wins, then the smaller merit, then the smaller index
    (tensor ops only: used by a rank that holds no candidates and by the CPU test backend)."""
    merit = rec[:, 0]
    idx = rec[:, 1].contiguous().view(torch.int64)
    key = torch.where(torch.isnan(merit), torch.full_like(merit, -float("inf")), merit)
    tie = key == key.min()
    r = torch.argmin(torch.where(tie, idx, torch.full_like(idx, _INT64_MAX)))
    return rec[r]


def _better(av, ai, bv, bi):
    """np.argmin order: a NaN wins, then the smaller value, then the smaller index."""
    an, bn = av != av, bv != bv
    if an or bn:
        return (ai < bi) if (an and bn) else an
    if av != bv:
        return av < bv
    return ai < bi


def shard_offset(m_local, backend, group=None):
    """First global row index of this rank's block: ranks own contiguous blocks in rank order, so
    global index = offset + local index and the lowest-index tie rule stays global.  One all-gather of the row counts
    (and one host read); callers with equal or otherwise known blocks pass the offset to ``sharded_select`` instead."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=backend.device)
    counts[rank] = m_local
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return int(counts[:rank].sum().item())


def sharded_select(backend, cand_local, Xdist, alias, weights, num_pts, dtol=1e-3, group=None, offset=None):
    """Run the selection loop over all ranks' shards.  Every rank returns the same
    ``(global_indices [num_pts], rows [num_pts, d])``.

    Per call: one MIN all-reduce of the statistics; per pick: one all-gather of the ranks' {merit, index, row} records
    -- every rank then finds the winner on its own device (`b200rbf_shard_apply`), so there is no broadcast and no host
    round trip -- and, except after the last pick, one MIN all-reduce of the refreshed distance statistics.  The picks
    are read back once, at the end."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    m_local = int(cand_local.shape[0])
    d = int(cand_local.shape[1])
    if offset is None:
        offset = shard_offset(m_local, backend, group)

    stats = backend.score(cand_local, Xdist, alias)  # {fmin, -fmax, dmin, -dmax}
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MIN, group=group)
    results = backend.new(num_pts * (2 + d), torch.float64).view(num_pts, 2 + d)
    for i in range(num_pts):
        rec = backend.best(float(weights[i]), dtol, stats, offset)
        if world > 1:
            gathered = backend.new(world * (2 + d), torch.float64)
            dist.all_gather_into_tensor(gathered, rec, group=group)
        else:
            gathered = rec
        last = i + 1 == num_pts
        # the owner poisons fvals[jj]; everyone folds the winning row into dmerit (candidate_srbf.py:50-55)
        new_stats = backend.apply(gathered, world, offset, not last, results[i])
        if not last:
            if world > 1:
                dist.all_reduce(new_stats, op=dist.ReduceOp.MIN, group=group)
            stats[2:] = new_stats[2:]  # fvals' min/max are computed once (candidate_srbf.py:40)
    host = results.cpu()  # the only device -> host read of the call
    backend.finish()
    picked = host[:, 1].contiguous().view(torch.int64).numpy().copy()
    return picked, host[:, 2:].numpy().copy()


def sharded_weighted_distance_merit(num_pts, surrogate, X, fX, cand_local, weights, Xpend=None, dtol=1e-3, group=None,
                                    offset=None):
    """weighted_distance_merit with ``cand`` split by rows over the ranks of ``group``.

    ``surrogate`` is this rank's GpuRBFInterpolant replica (same points on every rank).  ``cand_local`` is this
    rank's block of rows (numpy array or float64 CUDA tensor; it may be empty).  ``offset``: global index of its first
    row when the caller knows it (e.g. equal blocks: rank * rows), else it is derived with one all-gather.  Returns
    ``(new_points, global_indices)``.
    """
    X = np.asarray(X, dtype=np.float64)
    if Xpend is None:
        Xpend = np.empty([0, X.shape[1]])
    surrogate._fit()
    from .candidates import dist_points, same_points

    alias = same_points(X, surrogate.X)
    Xdist = dist_points(surrogate, X, Xpend, alias)
    idx, rows = sharded_select(GpuShardBackend(surrogate), cand_local, Xdist, alias, weights, num_pts, dtol, group, offset)
    return rows, idx
