"""Candidate scoring sharded over several GPUs: one process per GPU, torch.distributed (NCCL) for the plumbing.

weighted_distance_merit (candidate_srbf.py:8-57) couples candidates only through two min/max rescales and an
argmin, so each rank scores a contiguous block of candidate rows against its replica of the centers and
coefficients, and the ranks exchange only

  * ``{fmin, -fmax, dmin, -dmax}``   one all-reduce(MIN) after scoring, and ``{dmin, -dmax}`` again per extra pick
  * ``(merit, global index)`` pairs   one all-gather per pick; lowest merit wins, lowest index on ties
                                      (np.argmin's first-minimum rule, candidate_srbf.py:49)
  * the winning candidate row          one broadcast of d doubles from its owner per extra pick

No candidate data ever crosses NVLink.  The driver below is written against a small backend interface so the
host logic is testable on CPU with gloo; ``GpuShardBackend`` is the product backend (C ABI phase calls).
"""
import numpy as np


class GpuShardBackend:
    """Phase calls of include/b200rbf.h on this rank's handle; all buffers are CUDA tensors."""

    def __init__(self, surrogate):
        import torch

        self.torch = torch
        self.s = surrogate
        self.h = surrogate.handle
        self.device = torch.device("cuda", surrogate._device)

    def new(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype, device=self.device)

    def score(self, cand, Xdist, alias):
        t = self.torch
        self.h.set_stream(t.cuda.current_stream(self.device).cuda_stream)
        if isinstance(cand, np.ndarray):
            self.m = cand.shape[0]
            self.h.score(cand, self.s.lb, self.s.ub, Xdist, alias)
        else:  # CUDA tensor hand-off
            assert cand.is_cuda and cand.dtype == t.float64 and cand.is_contiguous()
            self.m = cand.shape[0]
            self.h.score(cand, self.s.lb, self.s.ub, Xdist, alias, m=self.m)
        stats = self.new(4, t.float64)
        self.h.shard_stats(stats)
        return stats

    def argmin(self, w, dtol, stats, offset):
        pair = self.new(2, self.torch.float64)  # {merit, int64 index} as raw 8-byte words
        self.h.shard_argmin(w, dtol, stats, offset, pair)
        return pair

    def take(self, local_idx):
        row = self.new(self.s.dim, self.torch.float64)
        self.h.shard_take(local_idx, row)
        return row

    def update(self, row):
        self.h.shard_update(row)
        stats = self.new(4, self.torch.float64)
        self.h.shard_stats(stats)
        return stats

    def finish(self):
        self.h.set_stream(None)


def _better(av, ai, bv, bi):
    """np.argmin order: a NaN wins, then the smaller value, then the smaller index."""
    an, bn = av != av, bv != bv
    if an or bn:
        return (ai < bi) if (an and bn) else an
    if av != bv:
        return av < bv
    return ai < bi


def sharded_select(backend, cand_local, Xdist, alias, weights, num_pts, dtol=1e-3, group=None):
    """Run the selection loop over all ranks' shards.  Every rank returns the same
    ``(global_indices [num_pts], rows [num_pts, d])``; rows are None on ranks that passed no host array."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    m_local = int(cand_local.shape[0])
    d = int(cand_local.shape[1])
    # contiguous row blocks: global index = offset + local index (keeps lowest-index tie breaking global)
    counts = torch.zeros(world, dtype=torch.int64, device=backend.device)
    counts[rank] = m_local
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    counts = counts.cpu().tolist()
    offsets = np.concatenate(([0], np.cumsum(counts)))
    offset = int(offsets[rank])

    stats = backend.score(cand_local, Xdist, alias)  # {fmin, -fmax, dmin, -dmax}
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MIN, group=group)
    picked, rows = [], []
    for i in range(num_pts):
        pair = backend.argmin(float(weights[i]), dtol, stats, offset)
        if world > 1:
            allp = torch.empty(2 * world, dtype=pair.dtype, device=pair.device)
            dist.all_gather_into_tensor(allp, pair, group=group)
        else:
            allp = pair
        host = allp.cpu().contiguous()
        vals = host[0::2].numpy()
        idxs = host.view(torch.int64)[1::2].numpy()
        bv, bi = float(vals[0]), int(idxs[0])
        for r in range(1, world):
            if _better(float(vals[r]), int(idxs[r]), bv, bi):
                bv, bi = float(vals[r]), int(idxs[r])
        owner = int(np.searchsorted(offsets, bi, side="right") - 1)
        picked.append(bi)
        last = i + 1 == num_pts
        # the owner poisons fvals[jj] and publishes the row; everyone folds it into dmerit
        if rank == owner:
            row = backend.take(bi - offset)
        else:
            row = backend.new(d, torch.float64)
        if world > 1:
            dist.broadcast(row, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        rows.append(row.cpu().numpy().copy())
        if not last:
            new_stats = backend.update(row)
            if world > 1:
                dist.all_reduce(new_stats, op=dist.ReduceOp.MIN, group=group)
            stats[2:] = new_stats[2:]  # fvals' min/max are computed once (candidate_srbf.py:40)
    backend.finish()
    return np.array(picked, dtype=np.int64), np.vstack(rows)


def sharded_weighted_distance_merit(num_pts, surrogate, X, fX, cand_local, weights, Xpend=None, dtol=1e-3, group=None):
    """weighted_distance_merit with ``cand`` split by rows over the ranks of ``group``.

    ``surrogate`` is this rank's GpuRBFInterpolant replica (same points on every rank).  ``cand_local`` is this
    rank's block of rows (numpy array or float64 CUDA tensor).  Returns ``(new_points, global_indices)``.
    """
    X = np.asarray(X, dtype=np.float64)
    if Xpend is None:
        Xpend = np.empty([0, X.shape[1]])
    Xdist = np.vstack((X, Xpend))
    surrogate._fit()
    alias = X.shape == surrogate.X.shape and np.array_equal(X, surrogate.X)
    idx, rows = sharded_select(GpuShardBackend(surrogate), cand_local, Xdist, alias, weights, num_pts, dtol, group)
    return rows, idx
