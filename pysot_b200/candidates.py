"""Candidate-search acquisition on the GPU behind the reference's function signatures.

  weighted_distance_merit   pySOT/auxiliary_problems/candidate_srbf.py:8-57
  candidate_srbf            pySOT/auxiliary_problems/candidate_srbf.py:60-121
  candidate_dycors          pySOT/auxiliary_problems/candidate_dycors.py:8-94
  candidate_uniform         pySOT/auxiliary_problems/candidate_uniform.py:7-53

Candidate GENERATION stays on the host and consumes the global numpy RNG and scipy's truncnorm in exactly the
reference's call order, so a seeded run proposes bit-identical candidate sets (north star: "host-generated
candidate sets").  SCORING and SELECTION (the m x n work) run in libb200rbf.so: one fused pass for surrogate
value + min distance + min/max statistics, then the merit / argmin / distance-update loop on the device.

``install()`` rebinds the names the reference's strategy modules bound at import time, so SRBFStrategy,
DYCORSStrategy and SOPStrategy use these functions unchanged.
"""
import numpy as np
import scipy.stats as stats

from .pareto import nd_sorting
from .rbf import GpuRBFInterpolant


def _round_vars(x, int_var, lb, ub):
    """pySOT/utils.py:68-93 (in place)."""
    if len(int_var) > 0:
        x[:, int_var] = np.round(x[:, int_var])
        for i in int_var:
            x[np.where(x[:, i] < lb[i]), i] = lb[i]
            x[np.where(x[:, i] > ub[i]), i] = ub[i]
    return x


def _is_cuda_tensor(a):
    return hasattr(a, "is_cuda") and a.is_cuda


def weighted_distance_merit(num_pts, surrogate, X, fX, cand, weights, Xpend=None, dtol=1e-3):
    """Pick ``num_pts`` rows of ``cand`` by the weighted-distance merit (candidate_srbf.py:8-57).

    ``cand`` is a host array or (tensor hand-off) a contiguous float64 CUDA tensor.  Returns a ``(num_pts, dim)``
    float64 numpy array of rows copied from ``cand``.
    """
    if not isinstance(surrogate, GpuRBFInterpolant):
        return _reference_or_raise("weighted_distance_merit", surrogate, num_pts=num_pts, surrogate=surrogate, X=X, fX=fX,
                                   cand=cand, weights=weights, Xpend=Xpend, dtol=dtol)
    if _is_cuda_tensor(cand):  # the selected rows come back with the indices (b200rbf_select_rows): no gather, one copy
        return merit_select_indices(num_pts, surrogate, X, cand, weights, Xpend, dtol, want_rows=True)[2]
    idx = merit_select_indices(num_pts, surrogate, X, cand, weights, Xpend, dtol)[0]
    return np.asarray(cand, dtype=np.float64)[idx, :].copy()


def same_points(X, Y):
    """``X`` equals ``Y`` element for element: the aliasing test of weighted_distance_merit."""
    return X.shape == Y.shape and bool(np.array_equal(X, Y))


def dist_points(surrogate, X, Xpend, alias):
    """``vstack((X, Xpend))`` of candidate_srbf.py:35 without copying X where that can be avoided: nothing pending -> X
    itself; X equal to the surrogate's points (``alias``) -> the pending rows are written behind them into the spare
    rows of the surrogate's own growth buffer (scratch until the next ``add_points``) and that prefix is handed over.
    At the bench shape (20 000 x 50) the copy was 1 ms of a 62 ms step."""
    k = Xpend.shape[0]
    if k == 0:
        return X
    if alias:
        buf = surrogate.__dict__.get("_grow")
        n = X.shape[0]
        if buf is not None and surrogate.X.base is buf[0] and surrogate.X.shape[0] == n and n + k <= buf[0].shape[0]:
            buf[0][n:n + k] = Xpend
            return buf[0][:n + k]
    return np.vstack((X, Xpend))


def _reference_or_raise(name, surrogate_obj, **kw):
    """A surrogate that is not a GpuRBFInterpolant (pySOT's GP / MARS / polynomial / stock RBF surrogates) keeps working
    after install(): the call goes to the reference function install() replaced.  Without install() there is nothing to
    delegate to, and the GPU path itself never falls back to the CPU."""
    for (modname, attr), fn in _saved.items():
        if attr == name:
            return fn(**kw)
    raise TypeError("pysot_b200.%s needs a GpuRBFInterpolant (there is no CPU fallback); got %r"
                    % (name, type(surrogate_obj).__name__))


def merit_select_indices(num_pts, surrogate, X, cand, weights, Xpend=None, dtol=1e-3, want_scores=False, want_rows=False):
    """The device part of weighted_distance_merit.  Returns (indices, merit values[, fvals, dmerit, stats]), or with
    ``want_rows`` (indices, merit values, cand[indices] as a host array)."""
    X = np.asarray(X, dtype=np.float64)
    dim = X.shape[1]
    if _is_cuda_tensor(cand):
        import torch

        assert cand.dtype == torch.float64 and cand.is_contiguous()
    else:
        cand = np.asarray(cand, dtype=np.float64)
    if cand.ndim != 2 or cand.shape[1] != dim:
        raise ValueError("XA and XB must have the same number of columns (i.e. feature dimension.)")  # cdist's message
    if Xpend is None:
        Xpend = np.empty([0, dim])  # candidate_srbf.py:33-34
    if X.shape[0] + Xpend.shape[0] == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")  # np.amin, :36
    surrogate._fit()  # predict() would do this at candidate_srbf.py:39
    # Under the strategies the surrogate's points are exactly X (surrogate_strategy.py:427-429); then one
    # distance serves both metrics.  extra_points (surrogate_strategy.py:243-247) break the aliasing.
    alias = same_points(X, surrogate.X)
    Xdist = dist_points(surrogate, X, Xpend, alias)
    group = getattr(surrogate, "device_group", None)
    if group is not None and not _is_cuda_tensor(cand) and not want_scores:  # one process, several devices
        return group.select(num_pts, X, cand, weights, Xpend, dtol)
    h = surrogate.handle
    with _torch_stream(h, cand):
        fv, dm, st = h.score(cand, surrogate.lb, surrogate.ub, Xdist, alias, want_scores=want_scores, m=cand.shape[0])
        if want_rows:
            return h.select(weights, num_pts, dtol, dim=dim)
        idx, merit = h.select(weights, num_pts, dtol)
    if want_scores:
        return idx, merit, fv, dm, st
    return idx, merit


class _torch_stream:
    """While a CUDA tensor is handed to the library, its kernels are enqueued on torch's current stream for that
    device: whatever produced the tensor (torch ops, ``generate_on_device``) is ordered before them, and whatever
    consumes the results afterwards (``.cpu()``, indexing) after them.  With a host array the handle keeps its own
    stream."""

    def __init__(self, handle, tensor):
        self._h = handle if _is_cuda_tensor(tensor) else None
        self._t = tensor

    def __enter__(self):
        if self._h is not None:
            import torch

            self._h.set_stream(torch.cuda.current_stream(self._t.device).cuda_stream)

    def __exit__(self, *exc):
        if self._h is not None:
            self._h.set_stream(None)
        return False


# ------------------------------------------------------------------------------------------ where candidates are drawn
_GENERATOR = {"mode": "host", "seed": 0, "calls": 0}


def set_candidate_generator(mode="host", seed=0):
    """``"host"`` (default): numpy / scipy draws in the reference's exact call order -- seeded runs reproduce the
    reference's candidate sets.  ``"device"``: candidates are drawn on the GPU (Philox, `b200rbf_generate`) and never
    touch the host; same distributions, different streams.  Every call uses a fresh sub-stream
    of ``seed``."""
    if mode not in ("host", "device"):
        raise ValueError("mode must be 'host' or 'device'")
    _GENERATOR.update(mode=mode, seed=int(seed), calls=0)


def generate_on_device(kind, opt_prob, surrogate, X, fX, prob_perturb=1.0, sampling_radius=0.2, subset=None,
                       num_cand=None, xbest=None, seed=None):
    """Draw the candidate set of candidate_{dycors,srbf,uniform} on the surrogate's GPU; returns a CUDA tensor."""
    import torch

    xbest, num_cand, subset = _prepare(opt_prob, X, fX, subset, num_cand, xbest)
    sigma = None if kind == "uniform" else _scale_factors(opt_prob, sampling_radius, subset)
    if seed is None:
        seed = (_GENERATOR["seed"] * 0x9E3779B97F4A7C15 + _GENERATOR["calls"] * 0xD1B54A32D192ED03 + 1) & 0xFFFFFFFFFFFFFFFF
        _GENERATOR["calls"] += 1
    cand = torch.empty((int(num_cand), opt_prob.dim), dtype=torch.float64, device=torch.device("cuda", surrogate._device))
    with _torch_stream(surrogate.handle, cand):  # ordered with the torch operations on the tensor, before and after
        surrogate.handle.generate(kind, int(num_cand), xbest, opt_prob.lb, opt_prob.ub, sigma, prob_perturb,
                                  np.asarray(subset, dtype=np.int32), opt_prob.int_var, seed, cand)
    return cand


def _prepare(opt_prob, X, fX, subset, num_cand, xbest=None):
    if xbest is None:
        xbest = np.copy(X[np.argmin(fX), :]).ravel()
    if num_cand is None:
        num_cand = 100 * opt_prob.dim
    if subset is None:
        subset = np.arange(0, opt_prob.dim)
    return xbest, num_cand, subset


def _scale_factors(opt_prob, sampling_radius, subset):
    """sigma per coordinate, at least 1 for integer variables (candidate_srbf.py:100-103)."""
    sf = sampling_radius * (opt_prob.ub - opt_prob.lb)
    ind = np.intersect1d(opt_prob.int_var, subset)
    if len(ind) > 0:
        sf[ind] = np.maximum(sf[ind], 1.0)
    return sf


# scipy draws a truncated normal by inversion: U = random_state.uniform(size), x = truncnorm._ppf(U, a, b), then
# x * scale + loc (rv_generic.rvs / rv_generic._rvs).  The reference calls it once per coordinate with scalar a, b, and
# _ppf re-evaluates the two quantities that depend on (a, b) only -- log Phi(a) and the log mass of [a, b] -- for every
# sample.  _truncnorm_rvs_batch makes the same draws for all coordinates in one pass with those two hoisted; every
# remaining operation is the element-wise one _ppf applies, so the variates (and the state of numpy's global stream)
# are bit-identical.  It leans on scipy internals, so the first use checks it against the public truncnorm.rvs and any
# difference (or a scipy without those internals) switches it off for the process.
_FAST_TRUNCNORM = {"ok": None}


def _truncnorm_from_uniform(q, a, b, loc, scale, counts):
    from scipy import special as sc
    from scipy.stats import _continuous_distns as cd

    left = a < 0
    lcdf = np.where(left, cd._norm_logcdf(np.where(left, a, 0.0)), cd._norm_logcdf(np.where(left, 0.0, -b)))
    lmass = cd._log_gauss_mass(a, b)
    rep_left = np.repeat(left, counts)
    with np.errstate(divide="ignore"):
        logq = np.log(q) if left.all() else np.where(rep_left, np.log(q), np.log1p(-q))
    x = sc.ndtri_exp(_log_sum_pairs(np.repeat(lcdf, counts), logq + np.repeat(lmass, counts), cd._log_sum))
    if not left.all():
        x = np.where(rep_left, x, -x)
    x *= np.repeat(scale, counts)
    x += np.repeat(loc, counts)
    return x


def _log_sum_pairs(p, q, scipy_log_sum):
    """log(exp(p) + exp(q)) with the arithmetic of scipy.special.logsumexp([p, q], axis=0) -- which is what
    truncnorm._ppf calls -- for the generic element: the maximum is split off, the other term enters as
    log1p(exp(lo - hi)), then + log(1) + hi.  Ties and non-finite elements take scipy's own routine."""
    hi, lo = np.maximum(p, q), np.minimum(p, q)
    with np.errstate(all="ignore"):
        lo -= hi
        out = np.log1p(np.exp(lo, out=lo), out=lo)
        out += 0.0
        out += hi
    # lo - hi == 0 is a tie; a non-finite p, q or result shows up as a non-finite `out` or hi
    odd = ~np.isfinite(out) | (p == q)
    if odd.any():
        out[odd] = scipy_log_sum(p[odd], q[odd])
    return out


def _fast_truncnorm_ok():
    if _FAST_TRUNCNORM["ok"] is None:
        try:
            a = np.array([-1.3, 0.2, -7.5, 3.0, -0.0, 1e-3, -40.0, 0.0, -175.0, 12.0, -1e-9, -3.5])
            b = np.array([2.1, 0.9, -6.0, 9.0, 4.0, 8.0, 40.0, 1e-6, 0.0, 175.0, 1e-9, 5.0])
            loc = np.array([0.4, -3.0, 10.0, 0.0, 1.0, -1.0, 0.0, 2.0, 20.0, -15.0, 0.5, 0.0])
            scale = np.array([0.2, 7.0, 1.0, 0.05, 3.0, 1.0, 0.5, 7.0, 0.2, 0.2, 1e3, 1.0])
            counts = np.array([370, 5, 640, 0, 190, 80, 300, 100, 200, 200, 100, 1000])
            rs = np.random.RandomState(12345)
            want = np.concatenate([stats.truncnorm.rvs(a=a[i], b=b[i], loc=loc[i], scale=scale[i], size=int(counts[i]),
                                                       random_state=rs) for i in range(len(a))])
            got = _truncnorm_from_uniform(np.random.RandomState(12345).uniform(size=int(counts.sum())), a, b, loc,
                                          scale, counts)
            _FAST_TRUNCNORM["ok"] = bool(np.array_equal(want, got))
        except Exception:  # scipy without these internals
            _FAST_TRUNCNORM["ok"] = False
    return _FAST_TRUNCNORM["ok"]


def _truncnorm_rvs_batch(a, b, loc, scale, counts):
    """Concatenation of ``truncnorm.rvs(a[i], b[i], loc[i], scale[i], size=counts[i])`` for i = 0, 1, ... drawn from
    numpy's global stream in that order."""
    a, b, loc, scale = (np.asarray(v, dtype=np.float64) for v in (a, b, loc, scale))
    counts = np.asarray(counts, dtype=np.int64)
    if not _fast_truncnorm_ok() or not (np.all(a < b) and np.all(scale > 0)):  # odd arguments: scipy's own checks
        return np.concatenate([stats.truncnorm.rvs(a=a[i], b=b[i], loc=loc[i], scale=scale[i], size=int(counts[i]))
                               for i in range(len(a))] + [np.empty(0)])
    return _truncnorm_from_uniform(np.random.uniform(size=int(counts.sum())), a, b, loc, scale, counts)


def generate_srbf(opt_prob, X, fX, sampling_radius=0.2, subset=None, num_cand=None):
    """Truncated-normal perturbation of every coordinate in ``subset`` around xbest (candidate_srbf.py:91-116)."""
    xbest, num_cand, subset = _prepare(opt_prob, X, fX, subset, num_cand)
    sf = _scale_factors(opt_prob, sampling_radius, subset)
    cand = np.multiply(np.ones((num_cand, opt_prob.dim)), xbest)
    sub = np.asarray(subset, dtype=np.int64)
    if len(sub) > 0:
        lo, hi, s, xb = opt_prob.lb[sub], opt_prob.ub[sub], sf[sub], xbest[sub]
        draws = _truncnorm_rvs_batch((lo - xb) / s, (hi - xb) / s, xb, s, np.full(len(sub), num_cand))
        cand[:, sub] = draws.reshape(len(sub), num_cand).T  # a repeated coordinate keeps its last draw, as in the loop
    return _round_vars(cand, opt_prob.int_var, opt_prob.lb, opt_prob.ub)


def generate_dycors(opt_prob, X, fX, prob_perturb, sampling_radius=0.2, subset=None, num_cand=None, xbest=None):
    """DYCORS perturbation (candidate_dycors.py:55-89): each coordinate of ``subset`` is perturbed with
    probability ``prob_perturb``; rows with no perturbed coordinate get one at ``randint(0, len(subset) - 1)``
    (upper bound exclusive, so never the last one -- reference behaviour, kept for seeded parity, :78).  The mask
    column is indexed by ``i`` while the value is written to ``subset[i]`` (:83-84), also kept."""
    xbest, num_cand, subset = _prepare(opt_prob, X, fX, subset, num_cand, xbest)
    sf = _scale_factors(opt_prob, sampling_radius, subset)
    if len(subset) == 1:
        ar = np.ones((num_cand, 1))
    else:
        ar = np.random.rand(num_cand, len(subset)) < prob_perturb
        ind = np.where(np.sum(ar, axis=1) == 0)[0]
        ar[ind, np.random.randint(0, len(subset) - 1, size=len(ind))] = 1
    cand = np.multiply(np.ones((num_cand, opt_prob.dim)), xbest)
    sub = np.asarray(subset, dtype=np.int64)
    if len(sub) > 0:
        # the reference's loop `for i in subset` indexes the mask column, the bounds and xbest with i and writes
        # column subset[i] (candidate_dycors.py:83-89); kept as is
        lo, hi, s, xb = opt_prob.lb[sub], opt_prob.ub[sub], sf[sub], xbest[sub]
        rows = [np.where(ar[:, i] == 1)[0] for i in sub]
        draws = _truncnorm_rvs_batch((lo - xb) / s, (hi - xb) / s, xb, s, [len(r) for r in rows])
        off = 0
        for i, ind in zip(sub, rows):
            cand[ind, subset[i]] = draws[off:off + len(ind)]
            off += len(ind)
    return _round_vars(cand, opt_prob.int_var, opt_prob.lb, opt_prob.ub)


def generate_uniform(opt_prob, X, fX, subset=None, num_cand=None):
    """Uniform candidates in the box on ``subset``, xbest elsewhere (candidate_uniform.py:34-48)."""
    xbest, num_cand, subset = _prepare(opt_prob, X, fX, subset, num_cand)
    cand = np.multiply(np.ones((num_cand, opt_prob.dim)), xbest)
    cand[:, subset] = np.random.uniform(opt_prob.lb[subset], opt_prob.ub[subset], (num_cand, len(subset)))
    return _round_vars(cand, opt_prob.int_var, opt_prob.lb, opt_prob.ub)


class _Prefit:
    """Runs the surrogate's lazy refit on a worker thread while the caller draws candidates on the host: the fit is one
    ctypes call (the GIL is released for its duration) and the host generators are numpy / scipy work of comparable
    length, so in host-generator mode the fit disappears behind the random numbers.  Nothing else touches the handle
    until ``join()``; an exception raised by the fit is re-raised there.  The RNG is only ever used by the caller's
    thread, so seeded runs are unaffected."""

    def __init__(self, surrogate):
        self._exc = None
        self._thread = None
        if isinstance(surrogate, GpuRBFInterpolant) and not surrogate.updated and surrogate.num_pts >= surrogate.ntail:
            import threading

            self._thread = threading.Thread(target=self._run, args=(surrogate,), daemon=True)
            self._thread.start()

    def _run(self, surrogate):
        try:
            surrogate._fit()
        except BaseException as e:  # noqa: BLE001 -- handed to the caller's thread
            self._exc = e

    def join(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
            if self._exc is not None:
                raise self._exc


def candidate_srbf(num_pts, opt_prob, surrogate, X, fX, weights, Xpend=None, sampling_radius=0.2, subset=None,
                   dtol=1e-3, num_cand=None):
    """candidate_srbf.py:60-121."""
    if not isinstance(surrogate, GpuRBFInterpolant):
        return _reference_or_raise("candidate_srbf", surrogate, num_pts=num_pts, opt_prob=opt_prob, surrogate=surrogate, X=X,
                                   fX=fX, weights=weights, Xpend=Xpend, sampling_radius=sampling_radius, subset=subset,
                                   dtol=dtol, num_cand=num_cand)
    if _GENERATOR["mode"] == "device":
        cand = generate_on_device("srbf", opt_prob, surrogate, X, fX, 1.0, sampling_radius, subset, num_cand)
    else:
        pre = _Prefit(surrogate)
        try:
            cand = generate_srbf(opt_prob, X, fX, sampling_radius, subset, num_cand)
        finally:
            pre.join()
    return weighted_distance_merit(num_pts=num_pts, surrogate=surrogate, X=X, fX=fX, Xpend=Xpend, cand=cand,
                                   dtol=dtol, weights=weights)


def candidate_dycors(num_pts, opt_prob, surrogate, X, fX, weights, prob_perturb, Xpend=None, sampling_radius=0.2,
                     subset=None, dtol=1e-3, num_cand=None, xbest=None):
    """candidate_dycors.py:8-94."""
    if not isinstance(surrogate, GpuRBFInterpolant):
        return _reference_or_raise("candidate_dycors", surrogate, num_pts=num_pts, opt_prob=opt_prob, surrogate=surrogate,
                                   X=X, fX=fX, weights=weights, prob_perturb=prob_perturb, Xpend=Xpend,
                                   sampling_radius=sampling_radius, subset=subset, dtol=dtol, num_cand=num_cand, xbest=xbest)
    if _GENERATOR["mode"] == "device":
        cand = generate_on_device("dycors", opt_prob, surrogate, X, fX, prob_perturb, sampling_radius, subset, num_cand,
                                  xbest)
    else:
        pre = _Prefit(surrogate)
        try:
            cand = generate_dycors(opt_prob, X, fX, prob_perturb, sampling_radius, subset, num_cand, xbest)
        finally:
            pre.join()
    return weighted_distance_merit(num_pts=num_pts, surrogate=surrogate, X=X, fX=fX, Xpend=Xpend, cand=cand,
                                   dtol=dtol, weights=weights)


def candidate_uniform(num_pts, opt_prob, surrogate, X, fX, weights, Xpend=None, subset=None, dtol=1e-3,
                      num_cand=None):
    """candidate_uniform.py:7-53."""
    if not isinstance(surrogate, GpuRBFInterpolant):
        return _reference_or_raise("candidate_uniform", surrogate, num_pts=num_pts, opt_prob=opt_prob, surrogate=surrogate,
                                   X=X, fX=fX, weights=weights, Xpend=Xpend, subset=subset, dtol=dtol, num_cand=num_cand)
    if _GENERATOR["mode"] == "device":
        cand = generate_on_device("uniform", opt_prob, surrogate, X, fX, 1.0, 0.2, subset, num_cand)
    else:
        pre = _Prefit(surrogate)
        try:
            cand = generate_uniform(opt_prob, X, fX, subset, num_cand)
        finally:
            pre.join()
    return weighted_distance_merit(num_pts=num_pts, surrogate=surrogate, X=X, fX=fX, Xpend=Xpend, cand=cand,
                                   dtol=dtol, weights=weights)


# ---------------------------------------------------------------------------------------------- install / uninstall
# (module, attribute) pairs bound at import time in the reference:
#   srbf_strategy.py:6, dycors_strategy.py:5, sop_strategy.py:7, auxiliary_problems/__init__.py:2-4,
#   candidate_dycors.py:5, candidate_uniform.py:4, candidate_srbf.py:8
_BINDINGS = [
    ("pySOT.strategy.srbf_strategy", "candidate_srbf"),
    ("pySOT.strategy.dycors_strategy", "candidate_dycors"),
    ("pySOT.strategy.sop_strategy", "candidate_dycors"),
    ("pySOT.auxiliary_problems", "candidate_srbf"),
    ("pySOT.auxiliary_problems", "candidate_dycors"),
    ("pySOT.auxiliary_problems", "candidate_uniform"),
    ("pySOT.auxiliary_problems", "weighted_distance_merit"),
    ("pySOT.auxiliary_problems.candidate_srbf", "weighted_distance_merit"),
    ("pySOT.auxiliary_problems.candidate_srbf", "candidate_srbf"),
    ("pySOT.auxiliary_problems.candidate_dycors", "weighted_distance_merit"),
    ("pySOT.auxiliary_problems.candidate_dycors", "candidate_dycors"),
    ("pySOT.auxiliary_problems.candidate_uniform", "weighted_distance_merit"),
    ("pySOT.auxiliary_problems.candidate_uniform", "candidate_uniform"),
    # host side of SOPStrategy: numpy non-dominated sorting with the reference's exact output (pareto.py)
    ("pySOT.strategy.sop_strategy", "nd_sorting"),
    ("pySOT.utils", "nd_sorting"),
]
_saved = {}


def install():
    """Rebind pySOT's candidate functions to the GPU versions (and SOP's ``nd_sorting`` to the numpy version with
    identical output).  Returns the list of names rebound."""
    import importlib

    mine = globals()
    done = []
    for modname, attr in _BINDINGS:
        try:
            mod = importlib.import_module(modname)
        except Exception:
            continue
        if not hasattr(mod, attr) or getattr(mod, attr) is mine[attr]:
            continue
        _saved[(modname, attr)] = getattr(mod, attr)
        setattr(mod, attr, mine[attr])
        done.append("%s.%s" % (modname, attr))
    return done


def uninstall():
    """Restore what ``install`` replaced."""
    import importlib

    for (modname, attr), fn in list(_saved.items()):
        setattr(importlib.import_module(modname), attr, fn)
        del _saved[(modname, attr)]
