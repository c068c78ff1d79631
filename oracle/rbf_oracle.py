"""numpy/scipy restatement of pySOT's RBF surrogate + candidate acquisition.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity: PINNED against
the live reference by ``tools/make_golden.py`` and ``tests/golden/*.npz``.

The arithmetic of the reference lives in third-party compiled code that is not
under ``/root/reference`` (versions unpinned by its ``setup.py:16``; this image:
numpy 2.3.5, scipy 1.18.1, OpenBLAS 0.3.30): ``scipy.spatial.distance.cdist``
(plain sequential ``sum((a-b)**2)`` then ``sqrt``), ``scipy.linalg.lu_factor``
(LAPACK dgetrf, partial pivoting), ``solve_triangular`` (dtrtrs), ``cholesky``
(dpotrf), numpy ufuncs and ``np.dot``.  This file calls the same routines in
the same order at the reference's call sites, so it reproduces the reference
bit for bit on the same machine; every function cites the lines it follows
(paths relative to ``/root/reference``).

Kernel / tail are plain strings here: kernel in {"cubic", "tps", "linear"},
tail in {"linear", "constant"}.
"""
import numpy as np
import scipy.linalg as sla
import scipy.spatial.distance as ssd
import scipy.stats as sst

EPS = np.finfo(float).eps

KERNEL_ORDER = {"cubic": 2, "tps": 2, "linear": 1}  # kernels/*_kernel.py:13-14 (order attr)
TAIL_DEGREE = {"linear": 1, "constant": 0}  # tails/*_tail.py (degree attr)


# --------------------------------------------------------------------------- helpers (pySOT/utils.py)
def to_unit_box(x, lb, ub):
    """utils.py:21-33."""
    return (x - lb) / (ub - lb)


def from_unit_box(x, lb, ub):
    """utils.py:36-48."""
    return lb + (ub - lb) * x


def unit_rescale(x):
    """utils.py:51-65 -- all-equal input maps to ones."""
    hi, lo = x.max(), x.min()
    if hi == lo:
        return np.ones(x.shape)
    return (x - lo) / (hi - lo)


def round_vars(x, int_var, lb, ub):
    """utils.py:68-93 -- in place on x."""
    if len(int_var) > 0:
        x[:, int_var] = np.round(x[:, int_var])
        for i in int_var:
            x[np.where(x[:, i] < lb[i]), i] = lb[i]
            x[np.where(x[:, i] > ub[i]), i] = ub[i]
    return x


# --------------------------------------------------------------------------- kernels and tails
def phi(kernel, r):
    """kernels/cubic_kernel.py:15-25, tps_kernel.py:18-29, linear_kernel.py:18-28.

    TPS clamps r < eps to eps IN PLACE on the caller's array (tps_kernel.py:28).
    """
    if kernel == "cubic":
        return r ** 3
    if kernel == "tps":
        r[r < EPS] = EPS
        return (r ** 2) * np.log(r)
    if kernel == "linear":
        return r
    raise ValueError(kernel)


def dphi(kernel, r):
    """kernels/cubic_kernel.py:27-37, tps_kernel.py:31-42, linear_kernel.py:30-40."""
    if kernel == "cubic":
        return 3 * r ** 2
    if kernel == "tps":
        r[r < EPS] = EPS
        return r * (1 + 2 * np.log(r))
    if kernel == "linear":
        return np.ones(r.shape)
    raise ValueError(kernel)


def tail_dim(tail, dim):
    """tails/linear_tail.py:17, constant_tail.py:16 (dim_tail)."""
    return dim + 1 if tail == "linear" else 1


def tail_eval(tail, U, dim):
    """tails/linear_tail.py:19-31, constant_tail.py:18-30."""
    U = np.atleast_2d(U)
    if U.shape[1] != dim:
        raise ValueError("Input has the wrong number of dimensions")
    ones = np.ones((U.shape[0], 1))
    return np.hstack((ones, U)) if tail == "linear" else ones


def tail_deriv(tail, u, dim):
    """tails/linear_tail.py:33-46, constant_tail.py:32-44."""
    u = np.atleast_2d(u)
    if u.shape[1] != dim:
        raise ValueError("Input has the wrong number of dimensions")
    if tail == "linear":
        return np.hstack((np.zeros((dim, 1)), np.eye(dim)))
    return np.zeros((dim, 1))


def median_capping(x):
    """surrogate/output_transformations/median_capping.py:4-17."""
    x = x.copy()
    med = np.median(x)
    x[x > med] = med
    return x


# --------------------------------------------------------------------------- RBF interpolant
class OracleRBF:
    """State + algorithms of surrogate/surrogate.py:20-74 and surrogate/rbf.py:11-215."""

    def __init__(self, dim, lb, ub, kernel="cubic", tail="linear", eta=1e-6, output_transformation=None):
        if KERNEL_ORDER[kernel] - 1 > TAIL_DEGREE[tail]:  # rbf.py:85-86
            raise ValueError("Kernel and tail mismatch")
        self.dim, self.lb, self.ub = dim, lb, ub
        self.kernel, self.tail, self.eta = kernel, tail, eta
        self.ntail = tail_dim(tail, dim)
        self.T = output_transformation if output_transformation is not None else (lambda v: v)
        self.reset()

    def reset(self):  # surrogate.py:42-48, rbf.py:89-95
        self.num_pts = 0
        self.X = np.empty([0, self.dim])
        self._X = np.empty([0, self.dim])
        self.fX = np.empty([0, 1])
        self.updated = False
        self.L = self.U = self.piv = self.c = None

    def add_points(self, xx, fx):  # surrogate.py:50-74
        xx = np.atleast_2d(xx)
        fx = np.asarray(fx, dtype=float)
        if fx.ndim == 0:
            fx = fx[None]
        if fx.ndim == 1:
            fx = fx[:, None]
        assert xx.shape[0] == fx.shape[0] and xx.shape[1] == self.dim
        self.X = np.vstack((self.X, xx))
        self._X = to_unit_box(self.X, self.lb, self.ub)  # ALL rows re-mapped, surrogate.py:71
        self.fX = np.vstack((self.fX, fx))
        self.num_pts += xx.shape[0]
        self.updated = False

    # ---- rbf.py:97-168
    def fit(self):
        if self.updated:
            return
        n, t = self.num_pts, self.ntail
        nact = n + t
        if self.c is None:
            assert n >= t  # rbf.py:111
            Xu = self._X[:n, :]
            Phi = phi(self.kernel, ssd.cdist(Xu, Xu)) + self.eta * np.eye(n)  # rbf.py:114-115
            P = tail_eval(self.tail, Xu, self.dim)
            A = np.vstack((np.hstack((np.zeros((t, t)), P.T)), np.hstack((P, Phi))))  # rbf.py:119-121
            LU, ipiv = sla.lu_factor(A)  # rbf.py:123
            self.L = np.tril(LU, -1) + np.eye(nact)
            self.U = np.triu(LU)
            perm = np.arange(nact)  # rbf.py:128-130: LAPACK ipiv -> permutation vector
            for i in range(nact):
                perm[i], perm[ipiv[i]] = perm[ipiv[i]], perm[i]
            self.piv = perm
        else:  # block-LU extension, rbf.py:132-161
            k = self.c.shape[0] - t
            nnew = n - k
            kact = t + k
            Xu, Xn = self._X[:n, :], self._X[k:n, :]
            D = ssd.cdist(Xu, Xn)
            Pnew = np.vstack((tail_eval(self.tail, Xn, self.dim).T, phi(self.kernel, D[:k, :])))
            Phinew = phi(self.kernel, D[k:, :]) + self.eta * np.eye(nnew)
            L21 = np.zeros((kact, nnew))
            U12 = np.zeros((kact, nnew))
            for i in range(nnew):
                L21[:, i] = sla.solve_triangular(a=self.U, b=Pnew[:kact, i], lower=False, trans="T")
                U12[:, i] = sla.solve_triangular(a=self.L, b=Pnew[self.piv[:kact], i], lower=True, trans="N")
            L21 = L21.T
            try:
                C = sla.cholesky(a=Phinew - np.dot(L21, U12), lower=True)
            except Exception:  # rbf.py:151-153: any failure -> full refit
                self.c = None
                return self.fit()
            self.piv = np.hstack((self.piv, np.arange(kact, nact)))
            self.L = np.hstack((np.vstack((self.L, L21)), np.vstack((np.zeros((kact, nnew)), C))))
            self.U = np.vstack((np.hstack((self.U, U12)), np.hstack((np.zeros((nnew, kact)), C.T))))
        rhs = np.vstack((np.zeros((t, 1)), self.T(self.fX.copy())))  # rbf.py:164-165
        y = sla.solve_triangular(a=self.L, b=rhs[self.piv], lower=True)
        self.c = sla.solve_triangular(a=self.U, b=y, lower=False)
        self.updated = True

    # ---- rbf.py:170-185
    def predict(self, xx):
        self.fit()
        u = to_unit_box(np.atleast_2d(xx), self.lb, self.ub)
        ds = ssd.cdist(u, self._X)
        t = self.ntail
        return np.dot(phi(self.kernel, ds), self.c[t : t + self.num_pts]) + np.dot(
            tail_eval(self.tail, u, self.dim), self.c[:t]
        )

    # ---- rbf.py:187-215
    def predict_deriv(self, xx):
        self.fit()
        u = to_unit_box(np.atleast_2d(xx), self.lb, self.ub)
        if u.shape[1] != self.dim:
            raise ValueError("Input has incorrect number of dimensions")
        ds = ssd.cdist(self._X, u)
        ds[ds < EPS] = EPS  # rbf.py:201
        t = self.ntail
        out = np.zeros((u.shape[0], self.dim))
        for i in range(u.shape[0]):
            x = np.atleast_2d(u[i, :])
            g = np.dot(tail_deriv(self.tail, x, self.dim), self.c[:t]).transpose()
            diff = -(self._X.copy())
            diff += x
            r = np.atleast_2d(ds[:, i]).T
            diff *= np.multiply(dphi(self.kernel, r), self.c[t:]) / r
            g += np.sum(diff, 0)
            out[i, :] = g
        return out / (self.ub - self.lb)


# --------------------------------------------------------------------------- acquisition
def weighted_distance_merit(num_pts, surrogate, X, fX, cand, weights, Xpend=None, dtol=1e-3, trace=None):
    """auxiliary_problems/candidate_srbf.py:8-57.

    ``surrogate`` only needs ``predict``.  ``trace`` (a dict) optionally
    receives the intermediate arrays for the parity tests.
    """
    dim = X.shape[1]
    if Xpend is None:
        Xpend = np.empty([0, dim])
    dmerit = np.amin(ssd.cdist(cand, np.vstack((X, Xpend))), axis=1, keepdims=True)  # :35-36
    fvals = surrogate.predict(cand)
    if trace is not None:
        trace["fvals_raw"] = fvals.copy()
        trace["dmerit0"] = dmerit.copy()
    fvals = unit_rescale(fvals)  # :40
    picked, merits = [], []
    new_points = np.ones((num_pts, dim))
    for i in range(num_pts):
        w = weights[i]
        merit = w * fvals + (1.0 - w) * (1.0 - unit_rescale(np.copy(dmerit)))  # :46
        merit[dmerit < dtol] = np.inf  # :48
        jj = np.argmin(merit)  # :49
        picked.append(int(jj))
        merits.append(float(merit[jj, 0]))
        if trace is not None:
            trace.setdefault("merit", []).append(merit.copy())
        fvals[jj] = np.inf  # :50
        new_points[i, :] = cand[jj, :].copy()
        dmerit = np.minimum(dmerit, ssd.cdist(cand, np.atleast_2d(new_points[i, :])))  # :54-55
    if trace is not None:
        trace["picked"] = np.array(picked, dtype=np.int64)
        trace["picked_merit"] = np.array(merits)
    return new_points


def _scale_factors(opt_prob, sampling_radius, subset):
    """candidate_srbf.py:100-103 == candidate_dycors.py:66-69."""
    sf = sampling_radius * (opt_prob.ub - opt_prob.lb)
    ind = np.intersect1d(opt_prob.int_var, subset)
    if len(ind) > 0:
        sf[ind] = np.maximum(sf[ind], 1.0)
    return sf


def gen_srbf(opt_prob, X, fX, sampling_radius=0.2, subset=None, num_cand=None):
    """Candidate generation half of candidate_srbf.py:91-116 (global numpy RNG, scipy truncnorm)."""
    xbest = np.copy(X[np.argmin(fX), :]).ravel()
    if num_cand is None:
        num_cand = 100 * opt_prob.dim
    if subset is None:
        subset = np.arange(0, opt_prob.dim)
    sf = _scale_factors(opt_prob, sampling_radius, subset)
    cand = np.multiply(np.ones((num_cand, opt_prob.dim)), xbest)
    for i in subset:
        lo, hi, s = opt_prob.lb[i], opt_prob.ub[i], sf[i]
        cand[:, i] = sst.truncnorm.rvs(a=(lo - xbest[i]) / s, b=(hi - xbest[i]) / s, loc=xbest[i], scale=s, size=num_cand)
    return round_vars(cand, opt_prob.int_var, opt_prob.lb, opt_prob.ub)


def gen_dycors(opt_prob, X, fX, prob_perturb, sampling_radius=0.2, subset=None, num_cand=None, xbest=None):
    """Candidate generation half of candidate_dycors.py:55-89.

    Keeps the reference's quirks: ``randint(0, len(subset) - 1)`` never forces the last
    coordinate (:78); the mask column is indexed by ``i`` but written to ``subset[i]`` (:83-84).
    """
    if xbest is None:
        xbest = np.copy(X[np.argmin(fX), :]).ravel()
    if num_cand is None:
        num_cand = 100 * opt_prob.dim
    if subset is None:
        subset = np.arange(0, opt_prob.dim)
    sf = _scale_factors(opt_prob, sampling_radius, subset)
    if len(subset) == 1:
        ar = np.ones((num_cand, 1))
    else:
        ar = np.random.rand(num_cand, len(subset)) < prob_perturb
        ind = np.where(np.sum(ar, axis=1) == 0)[0]
        ar[ind, np.random.randint(0, len(subset) - 1, size=len(ind))] = 1
    cand = np.multiply(np.ones((num_cand, opt_prob.dim)), xbest)
    for i in subset:
        lo, hi, s = opt_prob.lb[i], opt_prob.ub[i], sf[i]
        ind = np.where(ar[:, i] == 1)[0]
        cand[ind, subset[i]] = sst.truncnorm.rvs(
            a=(lo - xbest[i]) / s, b=(hi - xbest[i]) / s, loc=xbest[i], scale=s, size=len(ind)
        )
    return round_vars(cand, opt_prob.int_var, opt_prob.lb, opt_prob.ub)


def gen_uniform(opt_prob, X, fX, subset=None, num_cand=None):
    """Candidate generation half of candidate_uniform.py:34-48."""
    xbest = np.copy(X[np.argmin(fX), :]).ravel()
    if num_cand is None:
        num_cand = 100 * opt_prob.dim
    if subset is None:
        subset = np.arange(0, opt_prob.dim)
    cand = np.multiply(np.ones((num_cand, opt_prob.dim)), xbest)
    cand[:, subset] = np.random.uniform(opt_prob.lb[subset], opt_prob.ub[subset], (num_cand, len(subset)))
    return round_vars(cand, opt_prob.int_var, opt_prob.lb, opt_prob.ub)


def candidate_srbf(num_pts, opt_prob, surrogate, X, fX, weights, Xpend=None, sampling_radius=0.2, subset=None,
                   dtol=1e-3, num_cand=None):
    """candidate_srbf.py:60-121."""
    cand = gen_srbf(opt_prob, X, fX, sampling_radius, subset, num_cand)
    return weighted_distance_merit(num_pts, surrogate, X, fX, cand, weights, Xpend, dtol)


def candidate_dycors(num_pts, opt_prob, surrogate, X, fX, weights, prob_perturb, Xpend=None, sampling_radius=0.2,
                     subset=None, dtol=1e-3, num_cand=None, xbest=None):
    """candidate_dycors.py:8-94."""
    cand = gen_dycors(opt_prob, X, fX, prob_perturb, sampling_radius, subset, num_cand, xbest)
    return weighted_distance_merit(num_pts, surrogate, X, fX, cand, weights, Xpend, dtol)


def candidate_uniform(num_pts, opt_prob, surrogate, X, fX, weights, Xpend=None, subset=None, dtol=1e-3, num_cand=None):
    """candidate_uniform.py:7-53."""
    cand = gen_uniform(opt_prob, X, fX, subset, num_cand)
    return weighted_distance_merit(num_pts, surrogate, X, fX, cand, weights, Xpend, dtol)


# --------------------------------------------------------------------------- tiled scoring (CPU baseline driver)
def score_tiled(rbf, X, cand, Xpend=None, chunk=8192, threads=1):
    """fvals / dmerit of weighted_distance_merit computed over ``chunk``-row slices of ``cand``.

    The reference materialises the full m x n distance matrix twice (candidate_srbf.py:35,
    rbf.py:181); at m >= 1e5, n = 2e4 that does not fit.  This driver calls the same
    routines per slice (row results are independent of the slicing) and, with
    ``threads`` > 1, runs slices on a thread pool (cdist and the ufuncs release the GIL).
    """
    from concurrent.futures import ThreadPoolExecutor

    rbf.fit()
    P = np.vstack((X, Xpend)) if Xpend is not None else X
    m = cand.shape[0]
    fv = np.empty((m, 1))
    dm = np.empty((m, 1))

    def work(lo):
        hi = min(lo + chunk, m)
        dm[lo:hi] = np.amin(ssd.cdist(cand[lo:hi], P), axis=1, keepdims=True)
        fv[lo:hi] = rbf.predict(cand[lo:hi])

    starts = list(range(0, m, chunk))
    if threads <= 1:
        for lo in starts:
            work(lo)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, starts))
    return fv, dm


def select_from_scores(cand, fvals, dmerit, weights, num_pts, dtol=1e-3):
    """Selection loop of candidate_srbf.py:40-55 on precomputed fvals / dmerit (both (m,1))."""
    fvals = unit_rescale(fvals.copy())
    dmerit = dmerit.copy()
    picked = []
    for i in range(num_pts):
        w = weights[i]
        merit = w * fvals + (1.0 - w) * (1.0 - unit_rescale(np.copy(dmerit)))
        merit[dmerit < dtol] = np.inf
        jj = int(np.argmin(merit))
        picked.append(jj)
        fvals[jj] = np.inf
        dmerit = np.minimum(dmerit, ssd.cdist(cand, np.atleast_2d(cand[jj, :])))
    return np.array(picked, dtype=np.int64)
