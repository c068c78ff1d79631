/*
 * b200rbf.h -- C ABI of libb200rbf.so: pySOT's RBF-surrogate hot path on one B200 (sm_100a).
 *
 * Drop-in boundary.  The reference is pure Python; its "FFI" for this path is the set of
 * numpy/scipy calls made by pySOT/surrogate/rbf.py and pySOT/auxiliary_problems/candidate_srbf.py.
 * Each entry point below names the reference lines it replaces (paths relative to the pySOT tree).
 * The Python side (pysot_b200/_lib.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200RBF_E* code on failure; it never throws,
 *     aborts or exits.  b200rbf_last_error() returns a thread-local message for the last failure.
 *   - all floating-point data is IEEE fp64, matrices are C-contiguous row-major.
 *   - data pointers may be HOST or DEVICE pointers (told apart with cudaPointerGetAttributes); host
 *     inputs are staged through pinned buffers and overlapped with compute, outputs are caller
 *     allocated.  Pointers documented "dev" must be device memory (e.g. a torch tensor's data_ptr()).
 *   - a handle is bound to one device and is NOT thread-safe (pySOT drives the surrogate from the
 *     single POAP controller thread); calls are ordered on the handle's stream and synchronise before
 *     returning anything to host memory.
 *   - there is no CPU fallback: without a CUDA device b200rbf_create fails.
 */
#ifndef B200RBF_H
#define B200RBF_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200rbf_ctx* b200rbf_handle;

/* kernels: pySOT/surrogate/kernels/{cubic,tps,linear}_kernel.py ; tails: surrogate/tails/{linear,constant}_tail.py */
enum { B200RBF_KERNEL_CUBIC = 0, B200RBF_KERNEL_TPS = 1, B200RBF_KERNEL_LINEAR = 2 };
enum { B200RBF_TAIL_LINEAR = 0, B200RBF_TAIL_CONSTANT = 1 };

enum {
  B200RBF_OK = 0,
  B200RBF_EINVAL = -1,    /* bad argument (reference: ValueError / AssertionError) */
  B200RBF_ECUDA = -2,     /* CUDA runtime failure */
  B200RBF_ESTATE = -3,    /* call order: no model loaded / nothing scored */
  B200RBF_ESINGULAR = -4, /* exactly singular pivot in the saddle-point LU */
  B200RBF_ENOMEM = -5,
  B200RBF_EREFIT = -6     /* b200rbf_extend only: nothing extendable resident / budget used up / pivot failed: call b200rbf_fit */
};

/* options for b200rbf_set_option */
enum {
  /* 1: squared distances use separate multiply and add, sequentially over the coordinates -- bit-identical
   *    to scipy's cdist (rbf.py:181, candidate_srbf.py:35).  0 (default): fused multiply-add.           */
  B200RBF_OPT_EXACT_DIST = 1,
  /* rows per host->device chunk of the candidate pipeline for PAGEABLE host candidates (default 131072, rounded to
   *    whole waves of the score kernel); pinned host candidates use chunks that grow geometrically from one wave  */
  B200RBF_OPT_CHUNK_ROWS = 2,
  /* LU panel width (default 128; multiple of 32)                                                       */
  B200RBF_OPT_LU_PANEL = 3,
  /* 1: factor panel k+1 on a second, high-priority stream while the rest of trailing update k runs.  Default 0:
   *    measured to give no overlap (two GEMM CTAs fill an SM's registers), so the LU keeps cooperative launches  */
  B200RBF_OPT_LU_LOOKAHEAD = 4,
  /* DMMA GEMM variant: 2 or 3 = cp.async stages at BK 16, 4 = 2 stages at BK 32; 0 = auto (4; 2 under LU look-ahead) */
  B200RBF_OPT_GEMM_STAGES = 5,
  /* 1 (default): surrogate values use the norm expansion |u|^2 + |c|^2 - 2 u.c with the dot products on the FP64
   *    tensor path and a direct recomputation of near pairs (score_mma.cuh; 1.9x the rate of the direct kernel at
   *    d = 50, relative error of r^2 <= ~1e-12; square roots / logarithms built from the hardware seeds, <= 1 ulp).
   *    0: direct differences everywhere.  Ignored when EXACT_DIST is set.                                        */
  B200RBF_OPT_MMA_DIST = 6,
  /* solver of b200rbf_fit: 0 (default) auto = null-space Cholesky for cubic / TPS + linear tail with n >= 2 ntail + 32
   *    and dim <= 159, pivoted LU otherwise and whenever the Cholesky reports a non-positive pivot; 1 = always LU;
   *    2 = Cholesky or error                                                                                         */
  B200RBF_OPT_FIT_METHOD = 7,
  /* 1 (default): a Cholesky fit of n <= 8192 points pre-allocates room for 2 n (at least 1024) so that b200rbf_extend
   *    can append points in place; 0: never keep an extendable factorisation                                      */
  B200RBF_OPT_INCREMENTAL = 8,
  /* -1 (default) auto = dim 28 .. 32 and 36 .. 64 for calls of >= 8192 rows, dim 65 .. 128 always; 1 = always, 0 = never: the dot products of the norm expansion run on the INT8
   *    tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM): unit-box coordinates are 46-bit fixed-point numbers split
   *    into six base-256 digit planes, the digit products of equal weight accumulate exactly in int32 (score_i8.cu; error
   *    in r^2 below that of the FP64 norm expansion).  A call whose coordinates leave [0, 1] is re-run on the FP64 tensor
   *    path.  Ignored with EXACT_DIST, with MMA_DIST = 0 and for dim > 128.                                         */
  B200RBF_OPT_I8_DIST = 9
};

const char* b200rbf_last_error(void);
int b200rbf_version(void);

/* Lifetime.  device = CUDA ordinal.  The handle owns resident centers / coefficients / LU workspace /
 * candidate workspace until destroy. */
int b200rbf_create(int device, b200rbf_handle* out);
int b200rbf_destroy(b200rbf_handle h);
/* Stream the handle enqueues on.  use_own != 0: the handle's private non-blocking stream (the default).
 * use_own == 0: cuda_stream is taken literally, e.g. torch.cuda.current_stream().cuda_stream -- note that torch's
 * default stream IS the NULL (legacy default) stream, so NULL is a valid value here, not "unset". */
int b200rbf_set_stream(b200rbf_handle h, void* cuda_stream, int use_own);
int b200rbf_set_option(b200rbf_handle h, int option, long long value);

/* RBFInterpolant._fit, initial branch + coefficient solve (rbf.py:110-130, 164-168):
 *   A = [[0, P^T], [P, phi(cdist(Xu, Xu)) + eta I]],  c = A^-1 [0; rhs]
 * by LU with partial pivoting, or -- same solution, half the flops, no pivoting -- by a Cholesky factorisation of
 * the system projected on null(P^T) when the kernel is conditionally positive definite (B200RBF_OPT_FIT_METHOD).
 * Xu: n x d centers already mapped to the unit box (surrogate.py:71); rhs: n transformed values
 * (output_transformation(fX), rbf.py:164).  c_out: n + ntail coefficients, tail first (rbf.py:183-184),
 * ntail = d + 1 (linear tail) or 1 (constant).  Centers and coefficients stay resident for predict /
 * predict_deriv / score.  Requires n >= ntail (rbf.py:111) and kernel order - 1 <= tail degree (rbf.py:85).
 * fit_seconds (optional): device time of assembly + factorisation + solves (CUDA events). */
int b200rbf_fit(b200rbf_handle h, const double* Xu, const double* rhs, int n, int d, int kernel, int tail,
                double eta, double* c_out, double* fit_seconds);

/* RBFInterpolant._fit, incremental branch (rbf.py:132-161): the resident model was fitted (b200rbf_fit, Cholesky solver)
 * or extended for the first n_old rows of Xu, and rows n_old .. n are new points.  The reference extends its LU factors by
 * the new rows / columns and a Cholesky factorisation of their Schur complement, O(k n^2); here the null-space Cholesky
 * factor grows by one row per point (fit_extend.cu), pre-allocated -- no O(n^2) regrowth -- and resident on the device.
 * Xu: ALL n points (only the new rows are read); rhs: all n transformed values; rhs_changed != 0: the values of old
 * points differ from the previous call (e.g. median capping moved), so the right-hand side is re-projected.
 * Returns B200RBF_EREFIT -- and leaves the decision to the caller, exactly like the fallback at rbf.py:149-153 -- when
 * there is no extendable factorisation, the point count exceeds the pre-allocated capacity or 4 x the base size (the
 * null-space basis is re-orthogonalised by a full fit then), or a new pivot is not positive.  last_fit_method = 3. */
int b200rbf_extend(b200rbf_handle h, const double* Xu, const double* rhs, int n, int rhs_changed, double* c_out,
                   double* fit_seconds);

/* Make centers + coefficients resident without fitting (replica GPUs of the sharded scoring path, or a
 * model restored from a pickle).  c: n + ntail, tail first. */
int b200rbf_load(b200rbf_handle h, const double* Xu, const double* c, int n, int d, int kernel, int tail);

/* RBFInterpolant.predict (rbf.py:170-185).  x: m x d points in ORIGINAL coordinates; lb, ub: d host
 * doubles (to_unit_box is applied on the device, utils.py:33).  out: m values. */
int b200rbf_predict(b200rbf_handle h, const double* x, long long m, const double* lb, const double* ub, double* out);

/* RBFInterpolant.predict_deriv (rbf.py:187-215).  out: m x d gradients w.r.t. original coordinates. */
int b200rbf_predict_deriv(b200rbf_handle h, const double* x, long long m, const double* lb, const double* ub,
                          double* out);

/* First half of weighted_distance_merit (candidate_srbf.py:31-40): one fused pass computes
 *   fvals[i]  = surrogate.predict(cand[i])                        (unit-box metric, rbf.py:180-185)
 *   dmerit[i] = min_j || cand[i] - Xdist[j] ||_2                  (ORIGINAL metric, candidate_srbf.py:35-36)
 * and their global min / max (unit_rescale, utils.py:60-65).  Xdist = vstack(X, Xpend), ndist rows.
 * alias_centers != 0 promises that the first n rows of Xdist are the resident centers in original
 * coordinates (true under every pySOT strategy, surrogate_strategy.py:427-429); with an isotropic box the
 * kernel then derives both metrics from one distance.  cand / fvals / dmerit stay resident for select.
 * fvals_out / dmerit_out (optional, m each) and stats_out (optional, host, {fmin,fmax,dmin,dmax}). */
int b200rbf_score(b200rbf_handle h, const double* cand, long long m, const double* lb, const double* ub,
                  const double* Xdist, int ndist, int alias_centers, double* fvals_out, double* dmerit_out,
                  double* stats_out);

/* Second half of weighted_distance_merit (candidate_srbf.py:43-55) on the resident scores: for i < p
 *   merit = w_i * fhat + (1 - w_i) * (1 - dhat);  merit[dmerit < dtol] = inf;  jj = first argmin;
 *   fvals[jj] = inf;  dmerit = minimum(dmerit, || cand - cand[jj] ||).
 * idx_out: p selected row indices into cand; merit_out (optional): their merit values. */
int b200rbf_select(b200rbf_handle h, const double* weights, int p, double dtol, long long* idx_out,
                   double* merit_out);
/* Same, and also returns the selected rows, new_points = cand[idx] (candidate_srbf.py:51): rows_out (optional) p x d HOST
 * doubles, delivered by the call's one device-to-host copy -- with device-resident candidates the caller needs no
 * gather of its own. */
int b200rbf_select_rows(b200rbf_handle h, const double* weights, int p, double dtol, long long* idx_out,
                        double* merit_out, double* rows_out);

/* Candidate generation on the device (opt-in; candidate_dycors.py:73-86, candidate_srbf.py:108-113,
 * candidate_uniform.py:44-45, round_vars utils.py:68-93).  kind: 0 DYCORS, 1 SRBF, 2 uniform.  xbest, lb, ub, sigma
 * (= sampling_radius * (ub - lb), at least 1 for integer variables): d HOST doubles; subset: nsub coordinate indices
 * or NULL for all; int_mask: d bytes flagging integer variables or NULL; cand_dev: m x d DEVICE buffer.  Philox4x32-10
 * keyed by (seed, row, coordinate): same seed -> same candidates, but NOT numpy's / scipy's streams, so seeded runs do
 * not reproduce the reference's candidate sets (their distribution is identical).  The launch is asynchronous on the
 * handle's stream. */
int b200rbf_generate(b200rbf_handle h, int kind, long long m, int d, const double* xbest, const double* lb,
                     const double* ub, const double* sigma, double prob_perturb, const int* subset, int nsub,
                     const unsigned char* int_mask, unsigned long long seed, double* cand_dev);

/* ---- phase API for candidates sharded over several GPUs (one process and one handle per GPU).  The
 * caller reduces the tiny device buffers between phases with NCCL (torch.distributed):
 *   stats_dev  : 4 doubles {fmin, -fmax, dmin, -dmax}  -> all-reduce(MIN)
 *   pair_dev   : 2 x 8 bytes {double merit, int64 global_index} -> all-gather, lowest (merit, index) wins
 *   winner_dev : d doubles, the winning candidate row -> broadcast from its owner
 */
/* copy the local stats (after b200rbf_score or b200rbf_shard_update) into stats_dev in the layout above */
int b200rbf_shard_stats(b200rbf_handle h, double* stats_dev);
/* local merit + argmin using the (globally reduced) stats_dev; writes pair_dev; index = local + idx_offset */
int b200rbf_shard_argmin(b200rbf_handle h, double w, double dtol, const double* stats_dev, long long idx_offset,
                         void* pair_dev);
/* Device-side min-loc (no host round trip between the picks; what pysot_b200/sharded.py uses):
 *   best_dev   : 2 + d doubles {merit, int64 global index (raw bits), the candidate row behind it} -> all-gather
 *   gathered   : world x (2 + d) doubles, the ranks' best_dev records in rank order
 * shard_best  = local merit + argmin with the (globally reduced) stats_dev, writes best_dev.
 * shard_apply = picks the winner among the gathered records (lowest merit, then lowest global index, NaN first: np.argmin),
 *               copies the winning record to result_dev (2 + d doubles), marks the row taken on its owner
 *               (fvals[jj] = inf, candidate_srbf.py:50) and, when update != 0, folds the winner into dmerit
 *               (candidate_srbf.py:54-55) and writes this rank's new {fmin, -fmax, dmin, -dmax} to stats_dev. */
int b200rbf_shard_best(b200rbf_handle h, double w, double dtol, const double* stats_dev, long long idx_offset,
                       double* best_dev);
int b200rbf_shard_apply(b200rbf_handle h, const double* gathered_dev, int world, long long idx_offset, int update,
                        double* result_dev, double* stats_dev);
/* copy resident candidate row local_idx into winner_dev and mark fvals[local_idx] = inf (owner only) */
int b200rbf_shard_take(b200rbf_handle h, long long local_idx, double* winner_dev);
/* dmerit = minimum(dmerit, || cand - winner ||) on the local shard; refreshes the local dmin / dmax */
int b200rbf_shard_update(b200rbf_handle h, const double* winner_dev);

/* Measured FP64 ceilings of this device, used as roofline denominators by bench.py: a dependent-free
 * DFMA loop and an mma.sync m8n8k4 f64 (DMMA) loop, full grid, `ms` milliseconds each.
 * out[0] = DFMA TFLOP/s, out[1] = DMMA TFLOP/s, out[2] = combined TFLOP/s when every warp interleaves both
 * (tells whether the FP64 tensor op has its own datapath).  out must hold 3 doubles. */
int b200rbf_fp64_peaks(b200rbf_handle h, double ms, double* out);

/* Counters: number of kernels this handle launched since creation (bench.py's gpu_launches). */
long long b200rbf_launch_count(b200rbf_handle h);
/* Device time (ms, CUDA events on the handle's stream) of the fused score kernel(s) of the last b200rbf_score;
 * with host candidates it spans the chunk pipeline including the waits on the H2D copies. */
double b200rbf_last_score_ms(b200rbf_handle h);
/* Solver the last b200rbf_fit / b200rbf_extend used: 1 = pivoted LU, 2 = null-space Cholesky, 3 = row-append extension. */
int b200rbf_last_fit_method(b200rbf_handle h);
/* Kernel that computed the surrogate values of the last b200rbf_predict / b200rbf_score: 1 = direct differences
 * (score_kernel), 2 = FP64 tensor path (score_mma_kernel), 3 = INT8 tensor cores (score_i8_kernel), 4 = direct, d > 64. */
int b200rbf_last_score_kernel(b200rbf_handle h);

/* ---- test hooks: exercise the factorisation pieces in isolation (tests/ only, not the drop-in surface) ----
 * dbg_lu  : A_host is N x (N + nrhs) row-major; on return the leading N x N block holds L\U (unit lower L),
 *           the extra columns hold L^-1 P b, ipiv_host the LAPACK-style 0-based pivot rows, info the first
 *           exactly-zero pivot (1-based) or 0.                                  (scipy.linalg.lu_factor, rbf.py:123)
 * dbg_gemm: C (M x Nc) -= A (M x K) * B (K x Nc) on the FP64 tensor cores; K % 16 == 0. */
int b200rbf_dbg_lu(b200rbf_handle h, double* A_host, int N, int nrhs, int* ipiv_host, int* info);
int b200rbf_dbg_gemm(b200rbf_handle h, double* C_host, const double* A_host, const double* B_host, int M, int Nc,
                     int K, double* seconds);

#ifdef __cplusplus
}
#endif
#endif /* B200RBF_H */
