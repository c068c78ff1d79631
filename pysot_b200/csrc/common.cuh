// Shared declarations for libb200rbf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/b200rbf.h"

namespace b200rbf {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define B2_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      b200rbf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200RBF_ECUDA;                                                                       \
    }                                                                                             \
  } while (0)
#define B2_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)
#define B2_CHECK(cond, code, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      b200rbf::set_error(__VA_ARGS__);   \
      return (code);                     \
    }                                    \
  } while (0)

// Function attributes (the > 48 KB shared-memory opt-in) are PER DEVICE, and handles on different devices -- or a
// worker thread running a fit -- may reach the same launch site: one bit per device ordinal, set after the attribute
// calls succeeded.  Setting an attribute twice is harmless, so a benign race only repeats the call.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool pending(int device) const { return device >= 64 || !((mask.load(std::memory_order_acquire) >> device) & 1ull); }
  void done(int device) {
    if (device < 64) mask.fetch_or(1ull << device, std::memory_order_release);
  }
};

// ---------------------------------------------------------------- device buffers (grow-only)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return B200RBF_ENOMEM;
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMallocHost(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return B200RBF_ENOMEM;
    }
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// ---------------------------------------------------------------- constants
// tiling of the score kernels; the -D overrides exist for tools/k2_tune.cu (measured choices: DESIGN.md)
#ifndef B2_SCORE_THREADS
#define B2_SCORE_THREADS 128
#endif
#ifndef B2_SCORE_MINB
#define B2_SCORE_MINB 3
#endif
#ifndef B2_TILE_ROWS
#define B2_TILE_ROWS 32
#endif
#ifndef B2_STAGES
#define B2_STAGES 3
#endif
#ifndef B2_JT
#define B2_JT 4
#endif
constexpr int kMaxRegDim = 64;                  // candidate coordinates held in registers up to this (padded) dimension
constexpr int kScoreThreads = B2_SCORE_THREADS; // threads (= candidates) per CTA in the score kernels
constexpr int kScoreMinBlocks = B2_SCORE_MINB;  // __launch_bounds__ min CTAs per SM (register cap)
constexpr int kTileRows = B2_TILE_ROWS;         // center rows per shared-memory tile (TMA bulk copy granule)
constexpr int kStages = B2_STAGES;              // tile ring depth
constexpr int kJT = B2_JT;                      // centers processed together per thread (independent accumulators)
constexpr int kGenMaxDim = 192;                 // largest padded dimension of the shared-memory (generic) score kernel

inline int pad_dim(int d) { return (d + 1) & ~1; }                       // even: rows are multiples of 16 B
#ifndef B2_PAD_ROWS
#define B2_PAD_ROWS 32  // center rows are padded to multiples of this (duplicates of the last row); tuning builds may raise it
#endif
inline int pad_rows(int n) { return (n + B2_PAD_ROWS - 1) / B2_PAD_ROWS * B2_PAD_ROWS; }
inline int mma_stride(int d) {                                            // center row stride of the DMMA kernel:
  int s = (d + 3) & ~3;                                                   // >= 4 ceil(d/4), == 4 (mod 16) doubles so the
  while ((s & 15) != 4) s += 4;                                           // B-fragment LDS.64 are bank-conflict free
  return s;
}

// ---------------------------------------------------------------- the handle
struct Model {
  int n = 0, d = 0, dpad = 0, npad = 0, kernel = 0, tail = 0, ntail = 0;
  bool loaded = false;
  DevBuf centers;   // npad x dpad unit-box coordinates (pad rows = last row, pad dims = 0)
  DevBuf coef;      // npad RBF weights (pad = 0)
  DevBuf tailc;     // ntail tail coefficients (lambda_0, lambda_1..d)
  // operands of the DMMA score kernel (score_mma.cuh): centers with row stride ds and their squared norms
  DevBuf centers_x, cnorm;
  DevBuf centers_m;         // the same with (1, |c|^2) in slots d, d + 1 when `aug`
  bool aug = false;
  int ds = 0;
  // operands of the int8 tensor-core score kernel (score_i8.cu), built lazily on first use
  DevBuf centers_q, cnorm_q, coef_q, i8_flag;  // digit planes, norms of the quantised centers, weights padded to tiles,
  bool i8_ready = false, i8_bad = false;       // {centers out of [0,1], candidates out of [0,1] / fault}
  int i8_nk = 0, i8_ntiles = 0;
};

struct ScoreState {
  long long m = 0;          // resident candidates
  int d = 0;
  bool valid = false;
  const double* cand = nullptr;  // device pointer (own buffer or caller's)
  DevBuf cand_own;
  DevBuf fvals, dmerit;
  DevBuf parts;             // 4 x nparts per-CTA partial stats
  int nparts = 0;
  DevBuf pair_parts;        // per-CTA (merit, index) partials
  DevBuf stats;             // {fmin, fmax, dmin, dmax}
  DevBuf winner;            // dpad doubles
  DevBuf results;           // p x {int64 idx, double merit}
  DevBuf dist;              // padded distance point set
  DevBuf box;               // 2 x dpad : lb, (ub - lb)
};


// views into FitState::chol for one null-space Cholesky fit (P, W, Linv and the vectors are sized for `rows_cap` rows
// when the factorisation may be extended by fit_extend.cu)
struct CholWork {
  double *P, *Qneg, *W, *Acat, *Bcat, *T21, *Linv, *LinvT, *Tn, *G, *R1, *R2, *R, *S, *gpart, *v0, *v1, *v2, *q;
  double *Rinv, *fq, *bvec, *u, *mrow, *gbuf, *part;  // extension state: R^-1, Q0^T f_0, T^T f, L^-1 T^T f, scratch
  int n, n16, t, tp, ld;
};

// flags of the single-launch pipelined triangular solves (fit_extend.cu)
constexpr int kExtMaxBase = 8192;    // largest point count whose factorisation is kept extendable (capacity 2x)
constexpr int kExtGrowth = 4;        // refit from scratch once the point count exceeds this multiple of the base
constexpr int kTrsvMaxBlocks = 160;  // x 128 rows = 20 480
struct TrsvSync {
  unsigned ticket, done;
  unsigned err;  // a block gave up waiting for its predecessor
  int info;      // factorisation status: 0 ok, > 0 non-positive pivot near that row, < 0 tail QR failure

  unsigned flag[kTrsvMaxBlocks];
};

// what b200rbf_extend needs from the last null-space Cholesky fit (base points 0 .. n0) and the points appended since
struct ExtState {
  bool valid = false;
  int n0 = 0, N = 0, cap = 0, d = 0, kernel = 0;
  double eta = 0.0;
  CholWork w;
  DevBuf Xu;  // cap x d unit-box coordinates of all resident points
};

struct FitState {
  DevBuf A;      // N x ld saddle-point matrix, row-major, overwritten by L\U
  DevBuf piv;    // N int pivot rows
  DevBuf rhs;    // N
  DevBuf work;   // cooperative panel scratch
  DevBuf flags;
  DevBuf chol;   // workspace of the null-space Cholesky fit (fit_chol.cu)
  DevBuf sync;   // TrsvSync
  unsigned trsv_epoch = 0;
  ExtState ext;
};

}  // namespace b200rbf

struct b200rbf_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;  // cudaDevAttrMaxSharedMemoryPerBlockOptin
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_t0 = nullptr, ev_t1 = nullptr, ev_done = nullptr;
  cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;  // around the fused score kernel(s) of the last b200rbf_score
  double last_score_ms = -1.0;
  // LU look-ahead: panel k+1 factors on a high-priority stream while the rest of trailing update k runs
  cudaStream_t la_stream = nullptr;
  cudaEvent_t ev_la1 = nullptr, ev_la2 = nullptr;
  int fit_method = 0;  // 0 auto (null-space Cholesky where it applies, else pivoted LU), 1 LU, 2 Cholesky or fail
  int last_fit_method = 0;  // what the last b200rbf_fit actually used (1 LU, 2 Cholesky)
  int last_score_kernel = 0;  // value kernel of the last predict / score: 1 direct, 2 DMMA, 3 int8 tensor cores, 4 generic (d > 64)
  bool lu_lookahead = false;  // implemented and correct, but measured to give no overlap (see include/b200rbf.h)
  int gemm_stages = 4;          // DMMA GEMM variant: 2 / 3 = cp.async stages at BK 16, 4 = 2 stages at BK 32 (default)
  bool gemm_stages_auto = true; // drop to 2 while a look-ahead panel shares the SMs
  b200rbf::DevBuf tps_table;  // half-log table of the DMMA score kernel's thin-plate-spline epilogue (score_mma.cuh)
  b200rbf::PinBuf stage[2];
  b200rbf::PinBuf pin_small;
  long long launches = 0;
  bool exact_dist = false;
  bool mma_dist = true;  // norm-expansion distances on the FP64 tensor path (score_mma.cuh) unless exact_dist
  int i8_dist = -1;      // dot products of the norm expansion on the INT8 tensor cores (score_i8.cu): -1 auto (d >= 40), 0 off, 1 on
  long long cur_m = 0;   // rows of the predict / score call in progress (kernel choice)
  bool i8_veto = false;  // this call fell back (coordinates outside the unit box): the DMMA kernel serves it
  long long chunk_rows = 131072;  // target rows per H2D chunk; rounded to whole waves of the score kernel
  int lu_panel = 128;
  bool incremental = true;  // keep the Cholesky factorisation extendable by b200rbf_extend (pre-allocates 2x rows, n <= 8192)
  b200rbf::Model model;
  b200rbf::ScoreState score;
  b200rbf::FitState fit;
  b200rbf::DevBuf tmp_in, tmp_out, tmp_box;
  b200rbf::DevBuf gen_buf;  // parameters of the candidate generator
  b200rbf::PinBuf gen_pin;
};

namespace b200rbf {

// kernel launch entry points implemented in the .cu files -------------------------------------------
struct ScoreArgs {
  const double* cand;      // m x d (original coordinates), row-major, device
  long long m;
  int d;
  const double* lb;        // dpad (device) or nullptr => candidates are used as-is (no unit-box map)
  const double* range;     // dpad : ub - lb
  const double* pts;       // npad x DPAD point set (unit-box centers, or original-space distance set)
  const double* coef;      // npad weights (nullptr in distance-only mode)
  const double* tailc;     // ntail
  int npad;
  int tail;                // B200RBF_TAIL_*
  double dscale;           // distance-only / shared mode: factor from unit to original metric
  int min_mode;            // 0: none, 1: write dmerit = dscale*sqrt(min r2), 2: dmerit = min(dmerit, that)
  double* fvals;           // m (nullptr in distance-only mode)
  double* dmerit;          // m
  double* parts;           // 4 x nparts partial stats
  int nparts;
  int part_off;            // first partial slot of this launch
  int* occ_out;            // host only: non-null = do not launch, report resident CTAs per SM of the instantiation
};

// score kernel for kMaxRegDim < dpad <= kGenMaxDim (score_generic.cu) and its candidates-per-CTA
int launch_score_i8(b200rbf_ctx* h, const ScoreArgs& a);  // score_i8.cu
int i8_rows_per_cta();
int launch_score_generic(b200rbf_ctx* h, const ScoreArgs& a, int dpad, int kernel, bool exact, bool with_phi);
int score_generic_threads();
// parts: 4 rows {fmin, fmax, dmin, dmax} x `stride` slots, `count` of them valid; which: bit0 f rows, bit1 d rows
int launch_stats(b200rbf_ctx* h, const double* parts, int stride, int count, int which, double* stats);
int launch_merit_argmin(b200rbf_ctx* h, const double* fvals, const double* dmerit, long long m, const double* stats,
                        int stats_negated, double w, double dtol, long long idx_off, void* pair_parts, int* nblocks);
int launch_pick(b200rbf_ctx* h, const void* pair_parts, int nparts, const double* cand, int d, double* fvals,
                double* winner, int dpad, void* result_slot, void* pair_out);
int launch_best(b200rbf_ctx* h, const void* pair_parts, int nparts, const double* cand, int d, long long idx_off,
                double* best_out);
int launch_decide(b200rbf_ctx* h, const double* gathered, int world, int d, long long idx_off, long long m, double* fvals,
                  double* winner, int dpad, double* result);
int launch_take(b200rbf_ctx* h, const double* cand, int d, long long idx, double* fvals, double* winner, int dpad);
int launch_update(b200rbf_ctx* h, const double* cand, long long m, int d, const double* winner, double* dmerit,
                  double* parts, int stride, int* count_out, bool exact);
int launch_deriv(b200rbf_ctx* h, const double* x, long long m, const double* lb, const double* range, double* out,
                 double* scratch);
int run_fit(b200rbf_ctx* h, const double* Xu_dev, const double* rhs_dev, int n, int d, int kernel, int tail,
            double eta, double* c_dev, int* info);
int run_fp64_peaks(b200rbf_ctx* h, double ms, double* out);

}  // namespace b200rbf
