"""ctypes binding of libb200rbf.so (include/b200rbf.h) -- the only route from Python to the CUDA kernels.

There is no CPU fallback: if the shared library is missing or no B200 is visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RBF_LIB") or os.path.join(_HERE, "libb200rbf.so")  # override: A/B runs of two builds

KERNEL_CUBIC, KERNEL_TPS, KERNEL_LINEAR = 0, 1, 2
TAIL_LINEAR, TAIL_CONSTANT = 0, 1
OPT_EXACT_DIST, OPT_CHUNK_ROWS, OPT_LU_PANEL, OPT_LU_LOOKAHEAD, OPT_GEMM_STAGES, OPT_MMA_DIST, OPT_FIT_METHOD = 1, 2, 3, 4, 5, 6, 7
OPT_INCREMENTAL = 8
OPT_I8_DIST = 9
E_INVAL, E_CUDA, E_STATE, E_SINGULAR, E_NOMEM, E_REFIT = -1, -2, -3, -4, -5, -6

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/b200rbf.h one to one
PROTOTYPES = {
    "b200rbf_last_error": (C.c_char_p, []),
    "b200rbf_version": (C.c_int, []),
    "b200rbf_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "b200rbf_destroy": (C.c_int, [_vp]),
    "b200rbf_set_stream": (C.c_int, [_vp, _vp, C.c_int]),
    "b200rbf_set_option": (C.c_int, [_vp, C.c_int, C.c_longlong]),
    "b200rbf_fit": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _dp]),
    "b200rbf_extend": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _dp]),
    "b200rbf_load": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200rbf_predict": (C.c_int, [_vp, _vp, C.c_longlong, _vp, _vp, _vp]),
    "b200rbf_predict_deriv": (C.c_int, [_vp, _vp, C.c_longlong, _vp, _vp, _vp]),
    "b200rbf_score": (C.c_int, [_vp, _vp, C.c_longlong, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "b200rbf_select": (C.c_int, [_vp, _vp, C.c_int, C.c_double, _vp, _vp]),
    "b200rbf_select_rows": (C.c_int, [_vp, _vp, C.c_int, C.c_double, _vp, _vp, _vp]),
    "b200rbf_generate": (C.c_int, [_vp, C.c_int, C.c_longlong, C.c_int, _vp, _vp, _vp, _vp, C.c_double, _vp, C.c_int, _vp,
                                   C.c_ulonglong, _vp]),
    "b200rbf_shard_stats": (C.c_int, [_vp, _vp]),
    "b200rbf_shard_argmin": (C.c_int, [_vp, C.c_double, C.c_double, _vp, C.c_longlong, _vp]),
    "b200rbf_shard_best": (C.c_int, [_vp, C.c_double, C.c_double, _vp, C.c_longlong, _vp]),
    "b200rbf_shard_apply": (C.c_int, [_vp, _vp, C.c_int, C.c_longlong, C.c_int, _vp, _vp]),
    "b200rbf_shard_take": (C.c_int, [_vp, C.c_longlong, _vp]),
    "b200rbf_shard_update": (C.c_int, [_vp, _vp]),
    "b200rbf_fp64_peaks": (C.c_int, [_vp, C.c_double, _vp]),
    "b200rbf_launch_count": (C.c_longlong, [_vp]),
    "b200rbf_last_score_ms": (C.c_double, [_vp]),
    "b200rbf_last_fit_method": (C.c_int, [_vp]),
    "b200rbf_last_score_kernel": (C.c_int, [_vp]),
    "b200rbf_dbg_lu": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "b200rbf_dbg_gemm": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _dp]),
}

_lib = None


def load():
    """dlopen the library and declare every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "pysot_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pysot_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _raise(code):
    msg = load().b200rbf_last_error().decode("utf-8", "replace")
    if code == E_INVAL:
        raise ValueError(msg)
    if code == E_SINGULAR:
        raise np.linalg.LinAlgError(msg)
    if code == E_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError("libb200rbf: " + msg)


def check(code):
    if code != 0:
        _raise(code)


def f64(a):
    """C-contiguous float64 view/copy of a host array."""
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    """Raw address of a numpy array, a torch tensor (host or CUDA) or an int; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


class Handle:
    """Owns one b200rbf_handle (one CUDA device).  Not thread-safe, like the C handle."""

    def __init__(self, device=0):
        self._h = _vp()
        self.device = int(device)
        check(load().b200rbf_create(self.device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().b200rbf_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration
    def set_stream(self, stream_ptr):
        """stream_ptr: a cudaStream_t as int (0 is the legacy default stream, which is torch's default), or None
        to go back to the handle's private stream."""
        if stream_ptr is None:
            check(load().b200rbf_set_stream(self._h, None, 1))
        else:
            check(load().b200rbf_set_stream(self._h, int(stream_ptr), 0))

    def set_option(self, option, value):
        check(load().b200rbf_set_option(self._h, int(option), int(value)))

    @property
    def launches(self):
        return int(load().b200rbf_launch_count(self._h))

    @property
    def last_score_ms(self):
        return float(load().b200rbf_last_score_ms(self._h))

    @property
    def last_fit_method(self):
        return {1: "lu", 2: "cholesky", 3: "extension"}.get(int(load().b200rbf_last_fit_method(self._h)), "none")

    @property
    def last_score_kernel(self):
        return {1: "direct", 2: "dmma", 3: "int8", 4: "direct-generic"}.get(int(load().b200rbf_last_score_kernel(self._h)), "none")

    # ---- model
    def fit(self, Xu, rhs, kernel, tail, eta):
        Xu, rhs = f64(Xu), f64(rhs).ravel()
        n, d = Xu.shape
        nt = d + 1 if tail == TAIL_LINEAR else 1
        c = np.empty(n + nt)
        secs = C.c_double(0.0)
        check(load().b200rbf_fit(self._h, ptr(Xu), ptr(rhs), n, d, kernel, tail, float(eta), ptr(c), C.byref(secs)))
        return c, secs.value

    def extend(self, Xu, rhs, tail, rhs_changed=False):
        """Append the rows of Xu beyond the resident point count (rbf.py:132-161).  Returns (c, seconds), or None when the
        library asks for a full refit (nothing extendable resident, budget used up, pivot failed)."""
        Xu, rhs = f64(Xu), f64(rhs).ravel()
        n, d = Xu.shape
        nt = d + 1 if tail == TAIL_LINEAR else 1
        c = np.empty(n + nt)
        secs = C.c_double(0.0)
        code = load().b200rbf_extend(self._h, ptr(Xu), ptr(rhs), n, 1 if rhs_changed else 0, ptr(c), C.byref(secs))
        if code == E_REFIT:
            return None
        check(code)
        return c, secs.value

    def load_model(self, Xu, c, kernel, tail):
        Xu, c = f64(Xu), f64(c).ravel()
        n, d = Xu.shape
        check(load().b200rbf_load(self._h, ptr(Xu), ptr(c), n, d, kernel, tail))

    def predict(self, x, lb, ub, out=None, m=None):
        """x: (m, d) numpy array or CUDA tensor (then pass m and a CUDA tensor `out`)."""
        lb, ub = f64(lb), f64(ub)
        if isinstance(x, np.ndarray):
            x = f64(x)
            m = x.shape[0]
            if out is None:
                out = np.empty(m)
        check(load().b200rbf_predict(self._h, ptr(x), int(m), ptr(lb), ptr(ub), ptr(out)))
        return out

    def predict_deriv(self, x, lb, ub, out=None, m=None):
        lb, ub = f64(lb), f64(ub)
        if isinstance(x, np.ndarray):
            x = f64(x)
            m = x.shape[0]
            if out is None:
                out = np.empty((m, x.shape[1]))
        check(load().b200rbf_predict_deriv(self._h, ptr(x), int(m), ptr(lb), ptr(ub), ptr(out)))
        return out

    # ---- acquisition
    def score(self, cand, lb, ub, Xdist, alias_centers, want_scores=False, m=None):
        """cand: (m, d) numpy array or CUDA tensor.  Returns (fvals, dmerit, stats) -- arrays only if want_scores."""
        lb, ub, Xdist = f64(lb), f64(ub), f64(Xdist)
        if isinstance(cand, np.ndarray):
            cand = f64(cand)
            m = cand.shape[0]
        fv = np.empty(m) if want_scores else None
        dm = np.empty(m) if want_scores else None
        stats = np.empty(4)
        check(load().b200rbf_score(self._h, ptr(cand), int(m), ptr(lb), ptr(ub), ptr(Xdist), Xdist.shape[0],
                                   1 if alias_centers else 0, ptr(fv), ptr(dm), ptr(stats)))
        return fv, dm, stats

    def select(self, weights, num_pts, dtol, dim=None):
        w = f64(list(weights)[:num_pts])
        if w.shape[0] < num_pts:
            raise IndexError("list index out of range")  # weights[i] in candidate_srbf.py:45
        idx = np.empty(num_pts, dtype=np.int64)
        merit = np.empty(num_pts)
        if dim is not None:  # also the selected rows, cand[idx]
            rows = np.empty((num_pts, dim))
            check(load().b200rbf_select_rows(self._h, ptr(w), int(num_pts), float(dtol), ptr(idx), ptr(merit), ptr(rows)))
            return idx, merit, rows
        check(load().b200rbf_select(self._h, ptr(w), int(num_pts), float(dtol), ptr(idx), ptr(merit)))
        return idx, merit

    def generate(self, kind, m, xbest, lb, ub, sigma, prob_perturb, subset, int_var, seed, cand_dev):
        """Fill the CUDA tensor cand_dev (m, d) with candidates; kind in {"dycors", "srbf", "uniform"}."""
        xbest, lb, ub = f64(xbest), f64(lb), f64(ub)
        d = xbest.shape[0]
        sigma = f64(sigma) if sigma is not None else None
        sub = np.ascontiguousarray(subset, dtype=np.int32) if subset is not None else None
        mask = None
        if int_var is not None and len(int_var) > 0:
            mask = np.zeros(d, dtype=np.uint8)
            mask[np.asarray(int_var, dtype=int)] = 1
        check(load().b200rbf_generate(self._h, {"dycors": 0, "srbf": 1, "uniform": 2}[kind], int(m), d, ptr(xbest), ptr(lb),
                                      ptr(ub), ptr(sigma), float(prob_perturb), ptr(sub), 0 if sub is None else len(sub),
                                      ptr(mask), int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(cand_dev)))

    # ---- sharded phases (device pointers: torch CUDA tensors)
    def shard_stats(self, stats_dev):
        check(load().b200rbf_shard_stats(self._h, ptr(stats_dev)))

    def shard_argmin(self, w, dtol, stats_dev, idx_offset, pair_dev):
        check(load().b200rbf_shard_argmin(self._h, float(w), float(dtol), ptr(stats_dev), int(idx_offset), ptr(pair_dev)))

    def shard_best(self, w, dtol, stats_dev, idx_offset, best_dev):
        check(load().b200rbf_shard_best(self._h, float(w), float(dtol), ptr(stats_dev), int(idx_offset), ptr(best_dev)))

    def shard_apply(self, gathered_dev, world, idx_offset, update, result_dev, stats_dev):
        check(load().b200rbf_shard_apply(self._h, ptr(gathered_dev), int(world), int(idx_offset), 1 if update else 0,
                                         ptr(result_dev), ptr(stats_dev)))

    def shard_take(self, local_idx, winner_dev):
        check(load().b200rbf_shard_take(self._h, int(local_idx), ptr(winner_dev)))

    def shard_update(self, winner_dev):
        check(load().b200rbf_shard_update(self._h, ptr(winner_dev)))

    # ---- measurement / test hooks
    def fp64_peaks(self, ms=200.0):
        out = np.zeros(3)
        check(load().b200rbf_fp64_peaks(self._h, float(ms), ptr(out)))
        return {"dfma_tflops": float(out[0]), "dmma_tflops": float(out[1]), "dfma_plus_dmma_tflops": float(out[2])}

    def dbg_lu(self, A, nrhs=0):
        A = f64(A).copy()
        N = A.shape[0]
        assert A.shape[1] == N + nrhs
        ipiv = np.zeros(N, dtype=np.int32)
        info = np.zeros(1, dtype=np.int32)
        check(load().b200rbf_dbg_lu(self._h, ptr(A), N, nrhs, ptr(ipiv), ptr(info)))
        return A, ipiv, int(info[0])

    def dbg_gemm(self, Cm, A, B):
        Cm, A, B = f64(Cm).copy(), f64(A), f64(B)
        M, K = A.shape
        Nc = B.shape[1]
        secs = C.c_double(0.0)
        check(load().b200rbf_dbg_gemm(self._h, ptr(Cm), ptr(A), ptr(B), M, Nc, K, C.byref(secs)))
        return Cm, secs.value
