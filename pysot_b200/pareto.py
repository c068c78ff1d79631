"""Host-side non-dominated sorting for SOPStrategy.

``pySOT.utils.nd_sorting`` (utils.py:450-484) peels Pareto fronts with a pure-Python incremental skyline
(``nd_front`` / ``nd_add``, utils.py:392-447): O(N^2) - O(N^3) interpreter steps per call, called after every completed
evaluation (``sop_strategy.py:385-403``) -- 76 % of SOP's wall-clock in the reference.  Once the candidate search
runs on the GPU it is what is left of a proposal.  This is the same function on numpy arrays:

* domination is the reference's: strictly smaller in EVERY objective (``utils.py:389``; comparisons with NaN are
  false on both sides);
* front i is the set of remaining points no remaining point dominates -- unique, so the order in which the reference
  visits points does not matter;
* fronts are peeled until at least ``nmax`` points are ranked (``utils.py:473``: the test is made before a front, a
  front is always finished); unranked points keep ``inf``.

No N x N array is kept: domination counts are formed in row blocks and lowered front by front.
"""
import numpy as np

_BLOCK = 2048


def _dominates(Fa, Fb):
    """dom[i, j] = column i of Fa dominates column j of Fb."""
    dom = Fa[0][:, None] < Fb[0][None, :]
    for k in range(1, Fa.shape[0]):
        dom &= Fa[k][:, None] < Fb[k][None, :]
    return dom


def nd_sorting(F, nmax):
    """Non-domination rank (1, 2, ...) of each column of the (M, N) array ``F``; ``inf`` beyond the first fronts that
    cover ``nmax`` points.  Same return type as the reference: a float array of length N."""
    F = np.asarray(F, dtype=float)
    M, N = F.shape
    ranks = np.ones((N,), dtype=int) * float("inf")  # utils.py:468
    if N == 0:
        return ranks
    if M == 0:
        raise ValueError("nd_sorting needs at least one objective")
    ndom = np.zeros(N, dtype=np.int64)  # how many remaining points dominate column j
    for s in range(0, N, _BLOCK):
        ndom += _dominates(F[:, s:s + _BLOCK], F).sum(axis=0)
    remaining = np.ones(N, dtype=bool)
    count, i = 0, 1
    while count < nmax and remaining.any():
        front = remaining & (ndom == 0)
        if not front.any():  # cannot happen for a strict partial order; guards against an endless loop on bad input
            break
        ranks[front] = i
        count += int(front.sum())
        remaining &= ~front
        idx = np.nonzero(front)[0]
        for s in range(0, len(idx), _BLOCK):
            ndom -= _dominates(F[:, idx[s:s + _BLOCK]], F).sum(axis=0)
        i += 1
    return ranks
