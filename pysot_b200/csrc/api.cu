// C ABI of libb200rbf.so (include/b200rbf.h): argument checking, host<->device staging, stream ordering.
#include <stdarg.h>

#include <thread>

#include "score_mma.cuh"

namespace b200rbf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int deriv_scratch_doubles(const b200rbf_ctx* h, long long m, size_t* out);
int chol_fit_weights(b200rbf_ctx* h, const double* Xu_dev, const double* f_dev, int n, int d, int kernel, double eta,
                     CholWork* wk, int* info_dev, int cap_rows, double* c_out);
int chol_fit_tail(b200rbf_ctx* h, const CholWork& w, double* c_full_dev, double* tailc_out, int n_rows);
int chol_extend_weights(b200rbf_ctx* h, const double* f_dev, int n, bool rhs_changed, int* info_dev, double* c_out);
int ensure_fit_sync(b200rbf_ctx* h);
int run_generate(b200rbf_ctx* h, int kind, long long m, int d, const double* xbest, const double* lb, const double* ub,
                 const double* sigma, double prob_perturb, const int* subset, int nsub, const unsigned char* int_mask,
                 unsigned long long seed, double* cand_dev);
int run_lu_debug(b200rbf_ctx* h, double* A_dev, int ld, int N, int rcol0, int ncols, int* ipiv_dev, int* info_dev);
int run_gemm_debug(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M,
                   int Nc, int K);

// once per handle: the half-log table of the thin-plate-spline epilogues (score_mma.cuh, deriv_tile.cuh)
int ensure_tps_table(b200rbf_ctx* h) {
  if (h->tps_table.p != nullptr) return 0;
  double host[2 * kTpsLogEntries];
  tps_log_table_host(host);
  B2_TRY(h->tps_table.reserve(sizeof(host)));
  B2_CUDA(cudaMemcpy(h->tps_table.p, host, sizeof(host), cudaMemcpyHostToDevice));
  return 0;
}

namespace {

enum PtrKind { kPageable = 0, kPinned = 1, kDevice = 2 };
PtrKind ptr_kind(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return kPageable;
  }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return kDevice;
  if (at.type == cudaMemoryTypeHost) return kPinned;
  return kPageable;
}

// copy `bytes` from a host-or-device source into device memory, ordered on h->stream
int copy_in(b200rbf_ctx* h, void* dst_dev, const void* src, size_t bytes) {
  if (bytes == 0) return 0;
  const PtrKind k = ptr_kind(src);
  B2_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, k == kDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                          h->stream));
  return 0;
}
int copy_out(b200rbf_ctx* h, void* dst, const void* src_dev, size_t bytes) {
  if (bytes == 0) return 0;
  const PtrKind k = ptr_kind(dst);
  B2_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, k == kDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                          h->stream));
  return 0;
}

__global__ void pad_points_kernel(const double* __restrict__ src, int n, int d, double* __restrict__ dst, int npad,
                                  int dpad) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)npad * dpad) return;
  const int i = (int)(e / dpad), k = (int)(e - (long long)i * dpad);
  const int si = i < n ? i : n - 1;  // pad rows repeat the last point: no effect on a min, weight 0 in a sum
  dst[e] = k < d ? src[(size_t)si * d + k] : 0.0;
}
// The resident model in ONE launch (it was three): thread = padded center row i.
//   centers[i]   (dpad)  the row, pad rows repeat the last point (no effect on a min, weight 0 in a sum), pad dims 0
//   coef[i], tailc       the RBF weights (pad 0) and the tail coefficients, from c = [tail (t); weights (n)]
//   xrows[i] (ds), cnorm[i], aug[i]   operands of score_mma_kernel: the row with stride ds, its squared norm, and the
//                        augmented copy whose slots d and d + 1 carry 1 and |c|^2 (nullptr: not needed)
__global__ void install_kernel(const double* __restrict__ src, const double* __restrict__ c, int n, int d, int t,
                               double* __restrict__ centers, int npad, int dpad, double* __restrict__ coef,
                               double* __restrict__ tailc, double* __restrict__ xrows, int ds,
                               double* __restrict__ cnorm, double* __restrict__ aug) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < t) tailc[i] = c[i];
  if (i >= npad) return;
  coef[i] = i < n ? c[t + i] : 0.0;
  const int si = i < n ? i : n - 1;
  const double* row = src + (size_t)si * d;
  for (int k = 0; k < dpad; ++k) centers[(size_t)i * dpad + k] = k < d ? row[k] : 0.0;
  if (xrows == nullptr) return;
  double n2 = 0.0;
  for (int k = 0; k < ds; ++k) {
    const double v = k < d ? row[k] : 0.0;
    xrows[(size_t)i * ds + k] = v;
    if (aug != nullptr) aug[(size_t)i * ds + k] = v;
    n2 = fma(v, v, n2);
  }
  cnorm[i] = n2;
  if (aug != nullptr) {
    aug[(size_t)i * ds + d] = 1.0;
    aug[(size_t)i * ds + d + 1] = n2;
  }
}
__global__ void negate_stats_kernel(const double* __restrict__ s, double* __restrict__ out) {
  if (threadIdx.x == 0) {
    out[0] = s[0];
    out[1] = -s[1];
    out[2] = s[2];
    out[3] = -s[3];
  }
}

int check_model_args(int n, int d, int kernel, int tail) {
  B2_CHECK(n >= 1 && d >= 1, B200RBF_EINVAL, "n and d must be positive (n=%d, d=%d)", n, d);
  B2_CHECK(kernel >= 0 && kernel <= 2, B200RBF_EINVAL, "unknown kernel id %d", kernel);
  B2_CHECK(tail >= 0 && tail <= 1, B200RBF_EINVAL, "unknown tail id %d", tail);
  // rbf.py:85-86: kernel.order - 1 > tail.degree  ->  ValueError("Kernel and tail mismatch")
  const int order = (kernel == B200RBF_KERNEL_LINEAR) ? 1 : 2;
  const int degree = (tail == B200RBF_TAIL_LINEAR) ? 1 : 0;
  B2_CHECK(order - 1 <= degree, B200RBF_EINVAL, "Kernel and tail mismatch");
  B2_CHECK(pad_dim(d) <= kGenMaxDim, B200RBF_EINVAL, "dim %d > %d is not supported by this build", d, kGenMaxDim);
  return 0;
}

// make Xu (n x d, device) and c (n + t, device) the resident model
int install_model(b200rbf_ctx* h, const double* Xu_dev, const double* c_dev, int n, int d, int kernel, int tail) {
  Model& M = h->model;
  M.loaded = false;
  M.n = n;
  M.d = d;
  M.dpad = pad_dim(d);
  M.npad = pad_rows(n);
  M.kernel = kernel;
  M.tail = tail;
  M.ntail = (tail == B200RBF_TAIL_LINEAR) ? d + 1 : 1;
  B2_TRY(M.centers.reserve((size_t)M.npad * M.dpad * sizeof(double)));
  B2_TRY(M.coef.reserve((size_t)M.npad * sizeof(double)));
  B2_TRY(M.tailc.reserve((size_t)M.ntail * sizeof(double)));
  const bool mma = M.dpad <= kMaxRegDim;  // operands of the DMMA score kernel
  if (mma) {
    M.ds = mma_stride(d);
    B2_TRY(M.centers_x.reserve((size_t)M.npad * M.ds * sizeof(double)));
    B2_TRY(M.cnorm.reserve((size_t)M.npad * sizeof(double)));
    M.aug = (d % 4 == 1 || d % 4 == 2);  // two free slots in the last k-step of the DMMA kernel
    if (M.aug) B2_TRY(M.centers_m.reserve((size_t)M.npad * M.ds * sizeof(double)));
  }
  const int cnt = M.npad > M.ntail ? M.npad : M.ntail;
  install_kernel<<<(cnt + 127) / 128, 128, 0, h->stream>>>(
      Xu_dev, c_dev, n, d, M.ntail, M.centers.as<double>(), M.npad, M.dpad, M.coef.as<double>(), M.tailc.as<double>(),
      mma ? M.centers_x.as<double>() : nullptr, M.ds, mma ? M.cnorm.as<double>() : nullptr,
      (mma && M.aug) ? M.centers_m.as<double>() : nullptr);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  M.loaded = true;
  M.i8_ready = false;
  M.i8_bad = false;
  h->score.valid = false;
  return 0;
}

// lb / (ub - lb) -> device, padded to dpad (range pad = 1 so padded coordinates map to 0 / 1 = 0)
int upload_box(b200rbf_ctx* h, const double* lb, const double* ub, int d, int dpad, DevBuf& box, bool* isotropic,
               double* scale) {
  B2_TRY(h->pin_small.reserve(4096 * sizeof(double)));
  B2_CHECK(2 * dpad <= 4096, B200RBF_EINVAL, "dimension too large");
  // the pinned scratch may still be in flight from the previous call on this stream
  B2_CUDA(cudaStreamSynchronize(h->stream));
  double* hb = reinterpret_cast<double*>(h->pin_small.p);
  bool iso = true;
  for (int k = 0; k < dpad; ++k) {
    hb[k] = k < d ? lb[k] : 0.0;
    hb[dpad + k] = k < d ? ub[k] - lb[k] : 1.0;  // utils.py:33 denominator
    if (k < d && hb[dpad + k] != hb[dpad]) iso = false;
  }
  B2_TRY(box.reserve(2 * (size_t)dpad * sizeof(double)));
  B2_CUDA(cudaMemcpyAsync(box.p, hb, 2 * (size_t)dpad * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (isotropic) *isotropic = iso;
  if (scale) *scale = hb[dpad];
  return 0;
}

struct ScorePlan {
  bool with_phi = true;       // surrogate values
  bool shared_min = false;    // min distance to the centers derived from the unit-box distance (x dscale)
  double dscale = 1.0;
  const double* dist_pts = nullptr;  // padded original-space points for the distance-only pass (nullptr: none)
  int dist_npad = 0;
  int dist_mode = 0;          // min_mode of the distance-only pass (1 write, 2 combine)
};

// smallest candidates-per-CTA among the kernels that may serve this padded dimension (sizes the partial-stat slots)
int block_rows(int dpad) {
  if (dpad > kMaxRegDim) return score_generic_threads();
  return kMmaRowsPerCta < kScoreThreads ? kMmaRowsPerCta : kScoreThreads;
}
bool use_mma(const b200rbf_ctx* h, bool with_phi);
bool use_i8(const b200rbf_ctx* h, bool with_phi) {
  // auto (-1): where the int8 kernel beats the DMMA kernel -- one 32-byte k-step: d = 28 .. 32 (1.18x at 32); two and more:
  // from d = 36 (1.02x; 1.24x at 48, 1.54x at 64) -- and, above d = 64, the direct kernel that was the only one there (8x at
  // 96, 15x at 128); tools/i8_sweep.py, profiles/r02_i8_probes.txt
  // Few candidates (a strategy's 100 d per proposal) leave most SMs without a CTA of 128 rows and pay the per-CTA set-up
  // (digit planes, TMEM allocation): below 8192 rows the DMMA kernel stays (config #2 replay: 1.14 s against 1.33 s).
  const int dd = h->model.d;
  const bool want = h->i8_dist == 1 ||
                    (h->i8_dist < 0 && (dd > 64 || (h->cur_m >= 8192 && (dd >= 36 || (dd >= 28 && dd <= 32)))));
  return with_phi && want && h->mma_dist && !h->i8_veto && !h->model.i8_bad && !h->exact_dist && h->model.d <= 128;
}
// DMMA score kernel, few candidates: 4 (2) warps share a 16-row group and split the centers between them -- as long as
// that leaves at most ONE CTA per SM.  A call this small is latency-bound per warp, so two CTAs on an SM take twice as
// long per stage as one, and the kernel ends with the slowest SM: 3000 rows (188 groups) ran 46 us at split 4 (188 CTAs
// on 148 SMs) and 40 us at split 2 (94 CTAs).
int mma_split(const b200rbf_ctx* h) {
  const long long groups = (h->cur_m + 8 * kMmaMB - 1) / (8 * kMmaMB);
  const int per_cta = kMmaRowsPerCta / (8 * kMmaMB);  // row groups of a CTA at split 1
  for (int s = per_cta; s > 1; s >>= 1)
    if ((groups * s + per_cta - 1) / per_cta <= h->sm_count) return s;
  return 1;
}
// candidates per CTA of the kernel that computes surrogate values / of the distance-only kernel
int main_rows(const b200rbf_ctx* h) {
  if (use_i8(h, true)) return i8_rows_per_cta();
  if (h->model.dpad > kMaxRegDim) return score_generic_threads();
  return use_mma(h, true) ? kMmaRowsPerCta / mma_split(h) : kScoreThreads;
}
int dist_rows(const b200rbf_ctx* h) { return h->model.dpad > kMaxRegDim ? score_generic_threads() : kScoreThreads; }

bool use_mma(const b200rbf_ctx* h, bool with_phi) {
  return with_phi && h->mma_dist && !h->exact_dist && h->model.dpad <= kMaxRegDim;
}

int launch_score_mma(b200rbf_ctx* h, const ScoreArgs& a) {
  const Model& M = h->model;
  MmaArgs ma;
  ma.s = a;
  ma.aug = M.aug;
  ma.tps_table = nullptr;
  if (M.kernel == B200RBF_KERNEL_TPS) {
    B2_TRY(ensure_tps_table(h));
    ma.tps_table = h->tps_table.as<double2>();
  }
  ma.s.pts = M.aug ? M.centers_m.as<double>() : M.centers_x.as<double>();
  ma.cnorm = M.cnorm.as<double>();
  ma.ds = M.ds;
  ma.split = a.occ_out != nullptr ? 1 : mma_split(h);
  const int k4 = (M.d + 3) / 4;
  switch ((k4 - 1) / 4) {
    case 0: return launch_score_mma_group0(h, ma, k4, M.kernel);
    case 1: return launch_score_mma_group1(h, ma, k4, M.kernel);
    case 2: return launch_score_mma_group2(h, ma, k4, M.kernel);
    case 3: return launch_score_mma_group3(h, ma, k4, M.kernel);
  }
  set_error("internal: no DMMA score kernel for d = %d", M.d);
  return B200RBF_EINVAL;
}

int launch_score_any(b200rbf_ctx* h, const ScoreArgs& a, int dpad, int kernel, bool exact, bool with_phi) {
  const int which = use_i8(h, with_phi) ? 3 : use_mma(h, with_phi) ? 2 : dpad > kMaxRegDim ? 4 : 1;
  if (with_phi && a.occ_out == nullptr) h->last_score_kernel = which;
  if (which == 3) return launch_score_i8(h, a);
  if (which == 2) return launch_score_mma(h, a);
  if (dpad > kMaxRegDim) return launch_score_generic(h, a, dpad, kernel, exact, with_phi);
  switch ((dpad - 2) / 8) {
    case 0: return launch_score_group0(h, a, dpad, kernel, exact, with_phi);
    case 1: return launch_score_group1(h, a, dpad, kernel, exact, with_phi);
    case 2: return launch_score_group2(h, a, dpad, kernel, exact, with_phi);
    case 3: return launch_score_group3(h, a, dpad, kernel, exact, with_phi);
    case 4: return launch_score_group4(h, a, dpad, kernel, exact, with_phi);
    case 5: return launch_score_group5(h, a, dpad, kernel, exact, with_phi);
    case 6: return launch_score_group6(h, a, dpad, kernel, exact, with_phi);
    case 7: return launch_score_group7(h, a, dpad, kernel, exact, with_phi);
  }
  set_error("dim (padded %d) > %d is not supported by this build", dpad, kMaxRegDim);
  return B200RBF_EINVAL;
}

// run the fused kernels over rows [r0, r0 + rows) of the device-resident candidates
int score_rows(b200rbf_ctx* h, const ScorePlan& plan, const double* cand_dev, long long r0, long long rows,
               const double* box, double* fvals, double* dmerit, double* parts, int nparts, int off_f, int off_d) {
  const Model& M = h->model;
  ScoreArgs a;
  memset(&a, 0, sizeof(a));
  a.cand = cand_dev + r0 * M.d;
  a.m = rows;
  a.d = M.d;
  a.parts = parts;
  a.nparts = nparts;
  a.part_off = off_f;  // the value kernel's CTAs index the f rows (and the d rows when it also produces dmerit)
  if (plan.with_phi) {
    a.lb = box;
    a.range = box + M.dpad;
    a.pts = M.centers.as<double>();
    a.coef = M.coef.as<double>();
    a.tailc = M.tailc.as<double>();
    a.npad = M.npad;
    a.tail = M.tail;
    a.dscale = plan.dscale;
    a.min_mode = plan.shared_min ? 1 : 0;
    a.fvals = fvals + r0;
    a.dmerit = dmerit ? dmerit + r0 : nullptr;
    B2_TRY(launch_score_any(h, a, M.dpad, M.kernel, h->exact_dist, true));
  }
  if (plan.dist_pts != nullptr) {
    a.lb = nullptr;
    a.range = nullptr;
    a.pts = plan.dist_pts;
    a.coef = nullptr;
    a.tailc = nullptr;
    a.npad = plan.dist_npad;
    a.dscale = 1.0;
    a.min_mode = plan.dist_mode;
    a.part_off = off_d;  // its CTAs may cover a different number of candidates: separate slot numbering
    a.fvals = nullptr;
    a.dmerit = dmerit + r0;
    B2_TRY(launch_score_any(h, a, M.dpad, 0, h->exact_dist, false));
  }
  return 0;
}

// Pageable rows -> pinned staging buffer.  One host thread copies at 5 - 10 GB/s, about the rate at which the score
// kernel consumes candidates at the bench shape (7 GB/s), so on a slow host the staging bounded the whole call (84 ms
// for 419 MB against 60 ms of kernels); four threads leave it well under the kernel time.  Short-lived threads: a chunk
// is >= 8 MB here, the ~0.1 ms of creating them disappears in it, and no pool outlives the call.
void stage_copy(void* dst, const void* src, size_t bytes) {
  const size_t kMinParallel = (size_t)8 << 20;
  unsigned hw = std::thread::hardware_concurrency();
  const int nt = bytes < kMinParallel ? 1 : (hw >= 8 ? 4 : hw >= 4 ? 2 : 1);
  if (nt == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t slice = ((bytes / nt) + 4095) & ~(size_t)4095;
  std::thread th[3];
  int started = 0;
  for (int i = 1; i < nt; ++i) {
    const size_t off = slice * i;
    if (off >= bytes) break;
    const size_t len = bytes - off < slice ? bytes - off : slice;
    try {
      th[started] = std::thread([=] { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
      ++started;
    } catch (...) {  // no thread to be had: this slice is copied here
      memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len);
    }
  }
  memcpy(dst, src, slice < bytes ? slice : bytes);
  for (int i = 0; i < started; ++i) th[i].join();
}

// stream `m` candidate rows (host or device) through score_rows; host rows are staged chunk by chunk so the
// H2D copy of chunk i+1 overlaps the kernels of chunk i.  dst_dev receives the rows when they come from the host.
int score_pipeline(b200rbf_ctx* h, const ScorePlan& plan, const double* cand, long long m, double* dst_dev,
                   const double* box, double* fvals, double* dmerit, double* parts, int nparts,
                   const double** cand_dev_out, int* used_f, int* used_d) {
  const Model& M = h->model;
  const PtrKind kind = ptr_kind(cand);
  if (kind == kDevice) {
    const long long kMaxRows = (long long)block_rows(M.dpad) * 0x7fff0000LL;
    B2_CHECK(m <= kMaxRows, B200RBF_EINVAL, "too many candidates");
    B2_TRY(score_rows(h, plan, cand, 0, m, box, fvals, dmerit, parts, nparts, 0, 0));
    *cand_dev_out = cand;
    *used_f = (int)((m + main_rows(h) - 1) / main_rows(h));
    *used_d = plan.dist_pts ? (int)((m + dist_rows(h) - 1) / dist_rows(h)) : *used_f;
    return 0;
  }
  // chunk = a whole number of full waves of the score kernel (resident CTAs x candidates per CTA): consecutive
  // chunk kernels run back to back on one stream, so a chunk that ends in a mostly empty wave wastes the GPU
  long long chunk = h->chunk_rows, wave_rows = 0;
  {
    ScoreArgs q;
    memset(&q, 0, sizeof(q));
    int occ = 0;
    q.occ_out = &occ;
    B2_TRY(launch_score_any(h, q, M.dpad, M.kernel, h->exact_dist, plan.with_phi));
    const long long wave = (long long)h->sm_count * (occ > 0 ? occ : 1) * main_rows(h);
    long long waves = (chunk + wave / 2) / wave;
    if (waves < 1) waves = 1;
    chunk = waves * wave;
    wave_rows = wave;
  }
  const size_t row_bytes = (size_t)M.d * sizeof(double);
  if (kind == kPageable) {
    const long long cr = m < chunk ? m : chunk;
    B2_TRY(h->stage[0].reserve((size_t)cr * row_bytes));
    if (m > wave_rows) B2_TRY(h->stage[1].reserve((size_t)cr * row_bytes));  // more than one chunk: both buffers alternate
  }
  int off_f = 0, off_d = 0, slot = 0;
  bool used[2] = {false, false};
  // The chunks grow geometrically from one wave.  Only the first copy is exposed (one wave: 0.3 ms at the bench shape
  // instead of 0.9 ms), a copy runs ~10x faster than the kernel over the same rows so a chunk twice the size of the one
  // being scored still lands in time, and fewer launches mean fewer drained last waves.
  // Pageable source: the same, up to the size of the two staging buffers (`chunk` rows) -- what is exposed there is the
  // host memcpy into the staging buffer plus the copy of the FIRST chunk (2.6 + 0.5 ms at a fixed 64k-row chunk).
  long long cur = wave_rows;
  for (long long r0 = 0, rows = 0; r0 < m; r0 += rows, slot ^= 1) {
    rows = (m - r0 < cur) ? (m - r0) : cur;
    if (kind == kPinned) {
      if (m - r0 - rows < wave_rows) rows = m - r0;  // no sub-wave remainder launch
      cur *= 2;
    } else {
      if (m - r0 - rows < wave_rows && m - r0 <= chunk) rows = m - r0;
      // x1.5 (whole waves), not x2: staging the NEXT chunk (one host thread, ~10 GB/s) must not take longer than the
      // kernel over the current one (~7 GB/s of candidates at the bench shape), or the GPU waits for the host
      long long nxt = cur + (cur / (2 * wave_rows)) * wave_rows;
      if (nxt == cur) nxt = cur + wave_rows;
      cur = nxt < chunk ? nxt : chunk;
    }
    const void* src = cand + r0 * M.d;
    if (kind == kPageable) {
      if (used[slot]) B2_CUDA(cudaEventSynchronize(h->ev_copy[slot]));  // staging buffer free again
      stage_copy(h->stage[slot].p, src, (size_t)rows * row_bytes);
      src = h->stage[slot].p;
    }
    B2_CUDA(cudaMemcpyAsync(dst_dev + r0 * M.d, src, (size_t)rows * row_bytes, cudaMemcpyHostToDevice,
                            h->copy_stream));
    B2_CUDA(cudaEventRecord(h->ev_copy[slot], h->copy_stream));
    used[slot] = true;
    B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy[slot], 0));
    B2_TRY(score_rows(h, plan, dst_dev, r0, rows, box, fvals, dmerit, parts, nparts, off_f, off_d));
    off_f += (int)((rows + main_rows(h) - 1) / main_rows(h));
    off_d += (int)((rows + dist_rows(h) - 1) / dist_rows(h));
  }
  *cand_dev_out = dst_dev;
  *used_f = off_f;
  *used_d = plan.dist_pts ? off_d : off_f;
  return 0;
}

// ---- int8 score kernel: coordinates outside the unit box (or a fault in the tensor-core pipeline) are reported through
// two device flags; they are fetched with the call's final synchronisation and the call is repeated on the DMMA kernel
int* i8_flags_host(b200rbf_ctx* h) { return reinterpret_cast<int*>(h->pin_small.p) + 2 * 4000; }
int i8_queue_flags(b200rbf_ctx* h, bool used) {
  if (!used || !h->model.i8_ready) return 0;
  B2_CUDA(cudaMemcpyAsync(i8_flags_host(h), h->model.i8_flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}
int i8_after_sync(b200rbf_ctx* h, bool used, bool* retry) {
  *retry = false;
  if (!used || !h->model.i8_ready) return 0;
  const int* f = i8_flags_host(h);
  if (f[0] == 0 && f[1] == 0) return 0;
  B2_CHECK(f[1] != 2, B200RBF_ECUDA, "int8 score kernel: the tensor-core pipeline did not complete");
  if (f[0] != 0) h->model.i8_bad = true;  // a CENTER outside the unit box: this model stays on the DMMA kernel
  B2_CUDA(cudaMemsetAsync(h->model.i8_flag.as<int>() + 1, 0, sizeof(int), h->stream));
  *retry = true;
  return 0;
}

int count_parts(long long m, long long chunk, bool chunked, int br) {
  // upper bound on the per-CTA partial slots: one launch per chunk, each rounds its row count up to whole CTAs
  // (the pipeline may round `chunk` up to whole waves, which only lowers the number of launches)
  (void)chunk;
  (void)chunked;
  return (int)((m + br - 1) / br) + 4096;
}

}  // namespace
}  // namespace b200rbf

using namespace b200rbf;

extern "C" {

const char* b200rbf_last_error(void) { return g_err; }
int b200rbf_version(void) { return 100; }

int b200rbf_create(int device, b200rbf_handle* out) {
  B2_CHECK(out != nullptr, B200RBF_EINVAL, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); libb200rbf has no CPU fallback", cudaGetErrorString(e));
    cudaGetLastError();
    return B200RBF_ECUDA;
  }
  B2_CHECK(device >= 0 && device < count, B200RBF_EINVAL, "device %d out of range (%d devices)", device, count);
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  B2_CHECK(prop.major == 10, B200RBF_ECUDA, "libb200rbf is built for sm_100a only; device %d is sm_%d%d", device,
           prop.major, prop.minor);
  b200rbf_ctx* h = new (std::nothrow) b200rbf_ctx();
  B2_CHECK(h != nullptr, B200RBF_ENOMEM, "out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  auto init = [&]() -> int {
    B2_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    for (int i = 0; i < 2; ++i) B2_CUDA(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
    B2_CUDA(cudaEventCreate(&h->ev_t0));
    B2_CUDA(cudaEventCreate(&h->ev_t1));
    int lo = 0, hi = 0;  // numerically lower = higher priority
    B2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B2_CUDA(cudaStreamCreateWithPriority(&h->la_stream, cudaStreamNonBlocking, hi));
    B2_CUDA(cudaEventCreateWithFlags(&h->ev_la1, cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&h->ev_la2, cudaEventDisableTiming));
    B2_CUDA(cudaEventCreate(&h->ev_s0));
    B2_CUDA(cudaEventCreate(&h->ev_s1));
    B2_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    return 0;
  };
  const int rc = init();
  if (rc != 0) {  // a failure half way: release what exists (destroy tolerates the null members)
    b200rbf_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int b200rbf_destroy(b200rbf_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  Model& M = h->model;
  M.centers.release(); M.coef.release(); M.tailc.release();
  M.centers_x.release(); M.centers_m.release(); M.cnorm.release();
  M.centers_q.release(); M.cnorm_q.release(); M.coef_q.release(); M.i8_flag.release();
  ScoreState& S = h->score;
  S.cand_own.release(); S.fvals.release(); S.dmerit.release(); S.parts.release(); S.pair_parts.release();
  S.stats.release(); S.winner.release(); S.results.release(); S.dist.release(); S.box.release();
  FitState& F = h->fit;
  F.A.release(); F.piv.release(); F.rhs.release(); F.work.release(); F.flags.release(); F.chol.release();
  F.sync.release(); F.ext.Xu.release();
  h->tmp_in.release(); h->tmp_out.release(); h->tmp_box.release(); h->gen_buf.release(); h->gen_pin.release();
  h->tps_table.release();
  h->stage[0].release(); h->stage[1].release(); h->pin_small.release();
  cudaEvent_t evs[] = {h->ev_copy[0], h->ev_copy[1], h->ev_t0, h->ev_t1, h->ev_done, h->ev_s0, h->ev_s1, h->ev_la1, h->ev_la2};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  cudaStream_t streams[] = {h->la_stream, h->own_stream, h->copy_stream};
  for (cudaStream_t st : streams)
    if (st) {
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    }
  cudaGetLastError();
  delete h;
  return 0;
}

int b200rbf_set_stream(b200rbf_handle h, void* s, int use_own) {
  B2_CHECK(h != nullptr, B200RBF_EINVAL, "handle is NULL");
  h->stream = use_own ? h->own_stream : reinterpret_cast<cudaStream_t>(s);
  return 0;
}

int b200rbf_set_option(b200rbf_handle h, int option, long long value) {
  B2_CHECK(h != nullptr, B200RBF_EINVAL, "handle is NULL");
  switch (option) {
    case B200RBF_OPT_EXACT_DIST: h->exact_dist = value != 0; return 0;
    case B200RBF_OPT_MMA_DIST: h->mma_dist = value != 0; return 0;
    case B200RBF_OPT_I8_DIST:
      B2_CHECK(value >= -1 && value <= 1, B200RBF_EINVAL, "int8 mode must be -1 (auto), 0 (off) or 1 (on)");
      h->i8_dist = (int)value;
      return 0;
    case B200RBF_OPT_FIT_METHOD:
      B2_CHECK(value >= 0 && value <= 2, B200RBF_EINVAL, "fit method must be 0 (auto), 1 (LU) or 2 (Cholesky)");
      h->fit_method = (int)value;
      return 0;
    case B200RBF_OPT_CHUNK_ROWS:
      B2_CHECK(value >= 256, B200RBF_EINVAL, "chunk rows must be >= 256");
      h->chunk_rows = value;
      return 0;
    case B200RBF_OPT_LU_LOOKAHEAD: h->lu_lookahead = value != 0; return 0;
    case B200RBF_OPT_INCREMENTAL:
      h->incremental = value != 0;
      if (!h->incremental) h->fit.ext.valid = false;
      return 0;
    case B200RBF_OPT_GEMM_STAGES:
      B2_CHECK(value == 0 || (value >= 2 && value <= 4), B200RBF_EINVAL, "GEMM variant must be 0 (auto), 2, 3 or 4");
      h->gemm_stages_auto = value == 0;
      h->gemm_stages = value == 0 ? 4 : (int)value;
      return 0;
    case B200RBF_OPT_LU_PANEL:
      B2_CHECK(value >= 32 && value <= 128 && value % 32 == 0, B200RBF_EINVAL, "LU panel must be 32, 64, 96 or 128");
      h->lu_panel = (int)value;
      return 0;
  }
  set_error("unknown option %d", option);
  return B200RBF_EINVAL;
}

long long b200rbf_launch_count(b200rbf_handle h) { return h ? h->launches : 0; }
double b200rbf_last_score_ms(b200rbf_handle h) { return h ? h->last_score_ms : -1.0; }
int b200rbf_last_fit_method(b200rbf_handle h) { return h ? h->last_fit_method : 0; }
int b200rbf_last_score_kernel(b200rbf_handle h) { return h ? h->last_score_kernel : 0; }

int b200rbf_load(b200rbf_handle h, const double* Xu, const double* c, int n, int d, int kernel, int tail) {
  B2_CHECK(h && Xu && c, B200RBF_EINVAL, "NULL argument");
  B2_TRY(check_model_args(n, d, kernel, tail));
  B2_CUDA(cudaSetDevice(h->device));
  const int t = (tail == B200RBF_TAIL_LINEAR) ? d + 1 : 1;
  B2_TRY(h->tmp_in.reserve(((size_t)n * d + n + t) * sizeof(double)));
  double* Xd = h->tmp_in.as<double>();
  double* cd = Xd + (size_t)n * d;
  B2_TRY(copy_in(h, Xd, Xu, (size_t)n * d * sizeof(double)));
  B2_TRY(copy_in(h, cd, c, (size_t)(n + t) * sizeof(double)));
  h->fit.ext.valid = false;  // the resident factorisation (if any) no longer belongs to the resident model
  B2_TRY(install_model(h, Xd, cd, n, d, kernel, tail));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int b200rbf_fit(b200rbf_handle h, const double* Xu, const double* rhs, int n, int d, int kernel, int tail, double eta,
                double* c_out, double* fit_seconds) {
  B2_CHECK(h && Xu && rhs, B200RBF_EINVAL, "NULL argument");
  B2_TRY(check_model_args(n, d, kernel, tail));
  const int t = (tail == B200RBF_TAIL_LINEAR) ? d + 1 : 1;
  B2_CHECK(n >= t, B200RBF_EINVAL, "need at least ntail = %d points to fit, have %d (rbf.py:111)", t, n);
  B2_CUDA(cudaSetDevice(h->device));
  const int N = n + t;
  B2_TRY(h->tmp_in.reserve(((size_t)n * d + n + N) * sizeof(double)));
  double* Xd = h->tmp_in.as<double>();
  double* rd = Xd + (size_t)n * d;
  double* cd = rd + n;
  B2_TRY(h->pin_small.reserve(4096 * sizeof(double)));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  int* info_host = reinterpret_cast<int*>(h->pin_small.p);
  *info_host = 0;
  B2_TRY(copy_in(h, Xd, Xu, (size_t)n * d * sizeof(double)));
  B2_TRY(copy_in(h, rd, rhs, (size_t)n * sizeof(double)));
  B2_CUDA(cudaEventRecord(h->ev_t0, h->stream));
  // Null-space Cholesky (fit_chol.cu) where it applies: conditionally positive definite kernel + linear tail, and a
  // system large enough for the factorisation to matter.  Anything else, and any Cholesky failure, takes the pivoted LU.
  // its t x t helper kernels keep a (t rounded to 16)^2 matrix in shared memory: t <= 160 within the 227 KB opt-in limit
  const size_t tp16 = (size_t)((t + 15) & ~15);
  const bool chol_ok = tail == B200RBF_TAIL_LINEAR && kernel != B200RBF_KERNEL_LINEAR && n >= 2 * t + 32 &&
                       tp16 * tp16 * sizeof(double) <= (size_t)h->max_smem_optin;
  B2_CHECK(h->fit_method != 2 || chol_ok, B200RBF_EINVAL,
           "the Cholesky fit needs a cubic / TPS kernel with the linear tail, n >= 2 ntail + 32 and dim <= 159");
  bool done = false, installed = false;
  ExtState& E = h->fit.ext;
  E.valid = false;
  if (chol_ok && h->fit_method != 1) {  // faster than the LU from the smallest admissible size on (tools/fit_small.py)
    B2_TRY(ensure_fit_sync(h));
    TrsvSync* sy = h->fit.sync.as<TrsvSync>();
    int* info_dev = &sy->info;
    B2_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), h->stream));
    CholWork wk;
    // room for b200rbf_extend to append points in place: twice the rows (at least 1024), for the sizes where a refit
    // per added point is what a strategy would otherwise pay (rbf.py:132-161)
    int cap = 0;
    if (h->incremental && n <= kExtMaxBase) cap = 2 * n < 1024 ? 1024 : ((2 * n + 127) / 128) * 128;
    B2_TRY(chol_fit_weights(h, Xd, rd, n, d, kernel, eta, &wk, info_dev, cap, cd));  // cd = [0_t; weights]
    B2_TRY(chol_fit_tail(h, wk, cd, nullptr, n));  // cd = [lambda; weights]
    B2_TRY(install_model(h, Xd, cd, n, d, kernel, tail));
    if (cap > 0) {
      B2_TRY(E.Xu.reserve((size_t)cap * d * sizeof(double)));
      B2_CUDA(cudaMemcpyAsync(E.Xu.p, Xd, (size_t)n * d * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    B2_CUDA(cudaEventRecord(h->ev_t1, h->stream));
    unsigned* flags_host = reinterpret_cast<unsigned*>(info_host + 2);  // {err, info}
    B2_CUDA(cudaMemcpyAsync(flags_host, &sy->err, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    if (c_out) B2_TRY(copy_out(h, c_out, cd, (size_t)N * sizeof(double)));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    B2_CHECK(flags_host[0] == 0, B200RBF_ECUDA, "triangular solve: a block waited too long for its predecessor");
    *info_host = (int)flags_host[1];
    if (*info_host == 0) {
      done = true;
      installed = true;
      h->last_fit_method = 2;
      if (cap > 0) {
        E.valid = true;
        E.n0 = E.N = n;
        E.cap = cap;
        E.d = d;
        E.kernel = kernel;
        E.eta = eta;
        E.w = wk;
      }
    } else {
      h->model.loaded = false;
      B2_CHECK(h->fit_method != 2, B200RBF_ESINGULAR,
               "Cholesky fit failed (%s step, index %d): projected matrix not numerically positive definite",
               *info_host < 0 ? "tail QR" : "factorisation", *info_host < 0 ? -*info_host : *info_host);
      *info_host = 0;  // rbf.py:149-153: a failed Cholesky falls back to a fresh LU
    }
  }
  if (!done) {
    B2_TRY(run_fit(h, Xd, rd, n, d, kernel, tail, eta, cd, info_host));
    h->last_fit_method = 1;
  }
  if (!installed) {
    B2_TRY(install_model(h, Xd, cd, n, d, kernel, tail));
    B2_CUDA(cudaEventRecord(h->ev_t1, h->stream));
    if (c_out) B2_TRY(copy_out(h, c_out, cd, (size_t)N * sizeof(double)));
    B2_CUDA(cudaStreamSynchronize(h->stream));
  }
  if (fit_seconds) {
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
    *fit_seconds = ms * 1e-3;
  }
  if (*info_host != 0) {
    h->model.loaded = false;
    set_error("singular matrix: exactly zero pivot at diagonal %d of the saddle-point system", *info_host);
    return B200RBF_ESINGULAR;
  }
  return 0;
}

int b200rbf_extend(b200rbf_handle h, const double* Xu, const double* rhs, int n, int rhs_changed, double* c_out,
                   double* fit_seconds) {
  B2_CHECK(h && Xu && rhs, B200RBF_EINVAL, "NULL argument");
  ExtState& E = h->fit.ext;
  Model& M = h->model;
  // not extendable -> the caller does a full b200rbf_fit (which also re-orthogonalises the null-space basis)
  B2_CHECK(E.valid && M.loaded && M.n == E.N && M.d == E.d, B200RBF_EREFIT, "no extendable factorisation resident");
  B2_CHECK(n > E.N, B200RBF_EREFIT, "no new points (resident %d, given %d)", E.N, n);
  B2_CHECK(n <= E.cap && n <= kExtGrowth * E.n0, B200RBF_EREFIT,
           "%d points exceed the extension budget (capacity %d, %d x base %d)", n, E.cap, kExtGrowth, E.n0);
  B2_CUDA(cudaSetDevice(h->device));
  const int d = E.d, t = d + 1, N = n + t, n_old = E.N;
  B2_TRY(h->tmp_in.reserve(((size_t)n + N) * sizeof(double)));
  double* rd = h->tmp_in.as<double>();
  double* cd = rd + n;
  B2_TRY(h->pin_small.reserve(4096 * sizeof(double)));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  int* info_host = reinterpret_cast<int*>(h->pin_small.p);
  unsigned* sync_err = reinterpret_cast<unsigned*>(info_host + 1);
  *info_host = 0;
  double* Xall = E.Xu.as<double>();
  B2_TRY(copy_in(h, Xall + (size_t)n_old * d, Xu + (size_t)n_old * d, (size_t)(n - n_old) * d * sizeof(double)));
  B2_TRY(copy_in(h, rd, rhs, (size_t)n * sizeof(double)));
  B2_CUDA(cudaEventRecord(h->ev_t0, h->stream));
  B2_TRY(ensure_fit_sync(h));
  TrsvSync* sy = h->fit.sync.as<TrsvSync>();
  int* info_dev = &sy->info;
  B2_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), h->stream));
  E.valid = false;  // until this extension has been checked
  B2_TRY(chol_extend_weights(h, rd, n, rhs_changed != 0, info_dev, cd));  // cd = [0_t; weights]
  // tail exactly as in the full fit: lambda = R0^-1 (Q0^T f_0 - Wx^T c)
  B2_TRY(chol_fit_tail(h, E.w, cd, nullptr, n));
  B2_TRY(install_model(h, Xall, cd, n, d, E.kernel, B200RBF_TAIL_LINEAR));
  B2_CUDA(cudaEventRecord(h->ev_t1, h->stream));
  unsigned* flags_host = reinterpret_cast<unsigned*>(info_host + 2);  // {err, info}
  B2_CUDA(cudaMemcpyAsync(flags_host, &sy->err, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  if (c_out) B2_TRY(copy_out(h, c_out, cd, (size_t)N * sizeof(double)));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  *info_host = (int)flags_host[1];
  *sync_err = flags_host[0];
  if (fit_seconds) {
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
    *fit_seconds = ms * 1e-3;
  }
  B2_CHECK(*sync_err == 0, B200RBF_ECUDA, "triangular solve: a block waited too long for its predecessor");
  if (*info_host != 0) {  // rbf.py:149-153: the Cholesky of the new points' Schur complement failed -> full refit
    h->model.loaded = false;
    set_error("extension failed at point %d: not numerically positive definite", *info_host - 1);
    return B200RBF_EREFIT;
  }
  E.valid = true;
  h->last_fit_method = 3;
  return 0;
}

static int predict_impl(b200rbf_handle h, const double* x, long long m, const double* lb, const double* ub, double* out,
                        bool* retry) {
  *retry = false;
  B2_CHECK(h && x && lb && ub && out, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->model.loaded, B200RBF_ESTATE, "no model resident: call b200rbf_fit or b200rbf_load first");
  B2_CHECK(m >= 0, B200RBF_EINVAL, "negative row count");
  if (m == 0) return 0;
  B2_CUDA(cudaSetDevice(h->device));
  h->cur_m = m;
  const Model& M = h->model;
  B2_TRY(upload_box(h, lb, ub, M.d, M.dpad, h->tmp_box, nullptr, nullptr));
  const bool host_in = ptr_kind(x) != kDevice, host_out = ptr_kind(out) != kDevice;
  const int nparts = count_parts(m, h->chunk_rows, host_in, block_rows(M.dpad));
  size_t need = (size_t)4 * nparts * sizeof(double) + 64;
  const size_t off_f = need;
  if (host_out) need += (size_t)m * sizeof(double);
  B2_TRY(h->tmp_out.reserve(need));
  double* parts = h->tmp_out.as<double>();
  double* fv = host_out ? reinterpret_cast<double*>(reinterpret_cast<char*>(h->tmp_out.p) + off_f) : out;
  double* xdev = nullptr;
  if (host_in) {
    B2_TRY(h->tmp_in.reserve((size_t)m * M.d * sizeof(double)));
    xdev = h->tmp_in.as<double>();
  }
  ScorePlan plan;
  const double* used = nullptr;
  int used_f = 0, used_d = 0;
  B2_TRY(score_pipeline(h, plan, x, m, xdev, h->tmp_box.as<double>(), fv, nullptr, parts, nparts, &used, &used_f,
                        &used_d));
  if (host_out) B2_TRY(copy_out(h, out, fv, (size_t)m * sizeof(double)));
  const bool i8 = use_i8(h, true);
  B2_TRY(i8_queue_flags(h, i8));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  B2_TRY(i8_after_sync(h, i8, retry));
  return 0;
}

int b200rbf_predict(b200rbf_handle h, const double* x, long long m, const double* lb, const double* ub, double* out) {
  bool retry = false;
  int rc = predict_impl(h, x, m, lb, ub, out, &retry);
  if (rc == 0 && retry) {  // coordinates outside the unit box: the FP64 tensor path serves this call
    h->i8_veto = true;
    rc = predict_impl(h, x, m, lb, ub, out, &retry);
    h->i8_veto = false;
  }
  return rc;
}

int b200rbf_predict_deriv(b200rbf_handle h, const double* x, long long m, const double* lb, const double* ub,
                          double* out) {
  B2_CHECK(h && x && lb && ub && out, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->model.loaded, B200RBF_ESTATE, "no model resident: call b200rbf_fit or b200rbf_load first");
  B2_CHECK(m >= 0, B200RBF_EINVAL, "negative row count");
  if (m == 0) return 0;
  B2_CUDA(cudaSetDevice(h->device));
  const Model& M = h->model;
  B2_TRY(upload_box(h, lb, ub, M.d, M.dpad, h->tmp_box, nullptr, nullptr));
  const bool host_in = ptr_kind(x) != kDevice, host_out = ptr_kind(out) != kDevice;
  const double* xdev = x;
  if (host_in) {
    B2_TRY(h->tmp_in.reserve((size_t)m * M.d * sizeof(double)));
    B2_TRY(copy_in(h, h->tmp_in.p, x, (size_t)m * M.d * sizeof(double)));
    xdev = h->tmp_in.as<double>();
  }
  size_t scratch = 0;
  B2_TRY(deriv_scratch_doubles(h, m, &scratch));
  const size_t out_doubles = host_out ? (size_t)m * M.d : 0;
  B2_TRY(h->tmp_out.reserve((scratch + out_doubles) * sizeof(double)));
  double* sc = h->tmp_out.as<double>();
  double* od = host_out ? sc + scratch : out;
  B2_TRY(launch_deriv(h, xdev, m, h->tmp_box.as<double>(), h->tmp_box.as<double>() + M.dpad, od, sc));
  if (host_out) B2_TRY(copy_out(h, out, od, (size_t)m * M.d * sizeof(double)));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

static int score_impl(b200rbf_handle h, const double* cand, long long m, const double* lb, const double* ub,
                      const double* Xdist, int ndist, int alias_centers, double* fvals_out, double* dmerit_out,
                      double* stats_out, bool* retry) {
  *retry = false;
  B2_CHECK(h && cand && lb && ub && Xdist, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->model.loaded, B200RBF_ESTATE, "no model resident: call b200rbf_fit or b200rbf_load first");
  B2_CHECK(m >= 1, B200RBF_EINVAL, "need at least one candidate");
  B2_CHECK(ndist >= 1, B200RBF_EINVAL, "need at least one evaluated or pending point");
  B2_CUDA(cudaSetDevice(h->device));
  h->cur_m = m;
  const Model& M = h->model;
  ScoreState& S = h->score;
  S.valid = false;
  B2_CHECK(!alias_centers || ndist >= M.n, B200RBF_EINVAL, "alias_centers needs ndist >= n");
  bool iso = false;
  double scale = 1.0;
  B2_TRY(upload_box(h, lb, ub, M.d, M.dpad, S.box, &iso, &scale));

  ScorePlan plan;
  // EXACT_DIST promises cdist's bits in ORIGINAL coordinates: no shortcut through the unit-box distance then
  plan.shared_min = alias_centers && iso && !h->exact_dist;
  plan.dscale = plan.shared_min ? scale : 1.0;
  const int first = plan.shared_min ? M.n : 0;  // rows of Xdist that still need the original-space pass
  const int extra = ndist - first;
  if (extra > 0) {
    const int npd = pad_rows(extra);
    B2_TRY(S.dist.reserve(((size_t)npd * M.dpad + (size_t)extra * M.d) * sizeof(double)));
    double* padded = S.dist.as<double>();
    double* raw = padded + (size_t)npd * M.dpad;
    B2_TRY(copy_in(h, raw, Xdist + (size_t)first * M.d, (size_t)extra * M.d * sizeof(double)));
    const long long tot = (long long)npd * M.dpad;
    pad_points_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(raw, extra, M.d, padded, npd, M.dpad);
    h->launches++;
    B2_CUDA(cudaGetLastError());
    plan.dist_pts = padded;
    plan.dist_npad = npd;
    plan.dist_mode = plan.shared_min ? 2 : 1;
  }

  const bool host_in = ptr_kind(cand) != kDevice;
  const int nparts = count_parts(m, h->chunk_rows, host_in, block_rows(M.dpad));
  S.nparts = nparts > h->sm_count * 8 ? nparts : h->sm_count * 8;  // also holds the update kernel's partials
  B2_TRY(S.parts.reserve((size_t)4 * S.nparts * sizeof(double)));
  B2_TRY(S.fvals.reserve((size_t)m * sizeof(double)));
  B2_TRY(S.dmerit.reserve((size_t)m * sizeof(double)));
  B2_TRY(S.stats.reserve(64));
  B2_TRY(S.winner.reserve((size_t)M.dpad * sizeof(double)));
  B2_TRY(S.pair_parts.reserve((size_t)(h->sm_count * 8 + 8) * 16));
  double* own = nullptr;
  if (host_in) {
    B2_TRY(S.cand_own.reserve((size_t)m * M.d * sizeof(double)));
    own = S.cand_own.as<double>();
  }
  int used_f = 0, used_d = 0;
  B2_CUDA(cudaEventRecord(h->ev_s0, h->stream));
  B2_TRY(score_pipeline(h, plan, cand, m, own, S.box.as<double>(), S.fvals.as<double>(), S.dmerit.as<double>(),
                        S.parts.as<double>(), S.nparts, &S.cand, &used_f, &used_d));
  B2_CUDA(cudaEventRecord(h->ev_s1, h->stream));
  B2_TRY(launch_stats(h, S.parts.as<double>(), S.nparts, used_f, 1, S.stats.as<double>()));
  B2_TRY(launch_stats(h, S.parts.as<double>(), S.nparts, used_d, 2, S.stats.as<double>()));
  S.m = m;
  S.d = M.d;
  if (fvals_out) B2_TRY(copy_out(h, fvals_out, S.fvals.p, (size_t)m * sizeof(double)));
  if (dmerit_out) B2_TRY(copy_out(h, dmerit_out, S.dmerit.p, (size_t)m * sizeof(double)));
  if (stats_out) B2_TRY(copy_out(h, stats_out, S.stats.p, 4 * sizeof(double)));
  const bool i8 = use_i8(h, true);
  B2_TRY(i8_queue_flags(h, i8));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  B2_TRY(i8_after_sync(h, i8, retry));
  if (*retry) return 0;
  {
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev_s0, h->ev_s1) != cudaSuccess) cudaGetLastError();
    h->last_score_ms = ms;
  }
  S.valid = true;
  return 0;
}

int b200rbf_score(b200rbf_handle h, const double* cand, long long m, const double* lb, const double* ub,
                  const double* Xdist, int ndist, int alias_centers, double* fvals_out, double* dmerit_out,
                  double* stats_out) {
  bool retry = false;
  int rc = score_impl(h, cand, m, lb, ub, Xdist, ndist, alias_centers, fvals_out, dmerit_out, stats_out, &retry);
  if (rc == 0 && retry) {  // coordinates outside the unit box: the FP64 tensor path serves this call
    h->i8_veto = true;
    rc = score_impl(h, cand, m, lb, ub, Xdist, ndist, alias_centers, fvals_out, dmerit_out, stats_out, &retry);
    h->i8_veto = false;
  }
  return rc;
}

int b200rbf_select_rows(b200rbf_handle h, const double* weights, int p, double dtol, long long* idx_out,
                        double* merit_out, double* rows_out) {
  B2_CHECK(h && weights && idx_out, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  B2_CHECK(p >= 1, B200RBF_EINVAL, "num_pts must be >= 1");
  B2_CUDA(cudaSetDevice(h->device));
  ScoreState& S = h->score;
  const Model& M = h->model;
  const size_t slot = 16 + (size_t)S.d * 8;  // {merit, index, the selected row}
  B2_TRY(S.results.reserve((size_t)p * slot));
  B2_TRY(h->pin_small.reserve(((size_t)p * slot > 4096 * 8) ? (size_t)p * slot : 4096 * 8));
  char* res = reinterpret_cast<char*>(S.results.p);
  for (int i = 0; i < p; ++i) {
    int nb = 0;
    B2_TRY(launch_merit_argmin(h, S.fvals.as<double>(), S.dmerit.as<double>(), S.m, S.stats.as<double>(), 0,
                               weights[i], dtol, 0, S.pair_parts.p, &nb));
    B2_TRY(launch_pick(h, S.pair_parts.p, nb, S.cand, S.d, S.fvals.as<double>(), S.winner.as<double>(), M.dpad,
                       res + (size_t)i * slot, nullptr));
    if (i + 1 < p) {
      int cnt = 0;
      B2_TRY(launch_update(h, S.cand, S.m, S.d, S.winner.as<double>(), S.dmerit.as<double>(), S.parts.as<double>(),
                           S.nparts, &cnt, h->exact_dist));
      B2_TRY(launch_stats(h, S.parts.as<double>(), S.nparts, cnt, 2, S.stats.as<double>()));
    }
  }
  B2_CUDA(cudaMemcpyAsync(h->pin_small.p, res, (size_t)p * slot, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  const char* hp = reinterpret_cast<const char*>(h->pin_small.p);
  for (int i = 0; i < p; ++i) {
    double v;
    long long ix;
    memcpy(&v, hp + (size_t)i * slot, 8);
    memcpy(&ix, hp + (size_t)i * slot + 8, 8);
    idx_out[i] = ix;
    if (merit_out) merit_out[i] = v;
    if (rows_out) memcpy(rows_out + (size_t)i * S.d, hp + (size_t)i * slot + 16, (size_t)S.d * 8);
  }
  // the resident scores were consumed (fvals poisoned, dmerit updated): a second select needs a new score
  S.valid = false;
  return 0;
}

int b200rbf_select(b200rbf_handle h, const double* weights, int p, double dtol, long long* idx_out,
                   double* merit_out) {
  return b200rbf_select_rows(h, weights, p, dtol, idx_out, merit_out, nullptr);
}

int b200rbf_shard_stats(b200rbf_handle h, double* stats_dev) {
  B2_CHECK(h && stats_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  B2_CUDA(cudaSetDevice(h->device));
  negate_stats_kernel<<<1, 32, 0, h->stream>>>(h->score.stats.as<double>(), stats_dev);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int b200rbf_shard_argmin(b200rbf_handle h, double w, double dtol, const double* stats_dev, long long idx_offset,
                         void* pair_dev) {
  B2_CHECK(h && stats_dev && pair_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  B2_CUDA(cudaSetDevice(h->device));
  ScoreState& S = h->score;
  int nb = 0;
  B2_TRY(launch_merit_argmin(h, S.fvals.as<double>(), S.dmerit.as<double>(), S.m, stats_dev, 1, w, dtol, idx_offset,
                             S.pair_parts.p, &nb));
  B2_TRY(launch_pick(h, S.pair_parts.p, nb, nullptr, 0, nullptr, nullptr, 0, nullptr, pair_dev));
  return 0;
}

int b200rbf_shard_best(b200rbf_handle h, double w, double dtol, const double* stats_dev, long long idx_offset,
                       double* best_dev) {
  B2_CHECK(h && stats_dev && best_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  B2_CUDA(cudaSetDevice(h->device));
  ScoreState& S = h->score;
  int nb = 0;
  B2_TRY(launch_merit_argmin(h, S.fvals.as<double>(), S.dmerit.as<double>(), S.m, stats_dev, 1, w, dtol, idx_offset,
                             S.pair_parts.p, &nb));
  B2_TRY(launch_best(h, S.pair_parts.p, nb, S.cand, S.d, idx_offset, best_dev));
  return 0;
}

int b200rbf_shard_apply(b200rbf_handle h, const double* gathered_dev, int world, long long idx_offset, int update,
                        double* result_dev, double* stats_dev) {
  B2_CHECK(h && gathered_dev && result_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(world >= 1, B200RBF_EINVAL, "world must be >= 1");
  B2_CHECK(!update || stats_dev, B200RBF_EINVAL, "stats_dev is required with update");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  B2_CUDA(cudaSetDevice(h->device));
  ScoreState& S = h->score;
  const Model& M = h->model;
  B2_TRY(launch_decide(h, gathered_dev, world, S.d, idx_offset, S.m, S.fvals.as<double>(), S.winner.as<double>(), M.dpad,
                       result_dev));
  if (update) {
    int cnt = 0;
    B2_TRY(launch_update(h, S.cand, S.m, S.d, S.winner.as<double>(), S.dmerit.as<double>(), S.parts.as<double>(), S.nparts,
                         &cnt, h->exact_dist));
    B2_TRY(launch_stats(h, S.parts.as<double>(), S.nparts, cnt, 2, S.stats.as<double>()));
    negate_stats_kernel<<<1, 32, 0, h->stream>>>(S.stats.as<double>(), stats_dev);
    h->launches++;
    B2_CUDA(cudaGetLastError());
  }
  return 0;
}

int b200rbf_shard_take(b200rbf_handle h, long long local_idx, double* winner_dev) {
  B2_CHECK(h && winner_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  ScoreState& S = h->score;
  B2_CHECK(local_idx >= 0 && local_idx < S.m, B200RBF_EINVAL, "row %lld out of range", local_idx);
  B2_CUDA(cudaSetDevice(h->device));
  B2_TRY(launch_take(h, S.cand, S.d, local_idx, S.fvals.as<double>(), winner_dev, S.d));
  return 0;
}

int b200rbf_shard_update(b200rbf_handle h, const double* winner_dev) {
  B2_CHECK(h && winner_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(h->score.valid, B200RBF_ESTATE, "nothing scored: call b200rbf_score first");
  B2_CUDA(cudaSetDevice(h->device));
  ScoreState& S = h->score;
  int cnt = 0;
  B2_TRY(launch_update(h, S.cand, S.m, S.d, winner_dev, S.dmerit.as<double>(), S.parts.as<double>(), S.nparts, &cnt,
                       h->exact_dist));
  B2_TRY(launch_stats(h, S.parts.as<double>(), S.nparts, cnt, 2, S.stats.as<double>()));
  return 0;
}

int b200rbf_generate(b200rbf_handle h, int kind, long long m, int d, const double* xbest, const double* lb,
                     const double* ub, const double* sigma, double prob_perturb, const int* subset, int nsub,
                     const unsigned char* int_mask, unsigned long long seed, double* cand_dev) {
  B2_CHECK(h && xbest && lb && ub && cand_dev, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(kind >= 0 && kind <= 2, B200RBF_EINVAL, "unknown generator kind %d", kind);
  B2_CHECK(m >= 1 && d >= 1, B200RBF_EINVAL, "m and d must be positive");
  B2_CHECK(kind == 2 || sigma != nullptr, B200RBF_EINVAL, "sigma is required for the truncated-normal generators");
  B2_CHECK(ptr_kind(cand_dev) == kDevice, B200RBF_EINVAL, "cand_dev must be device memory");
  if (subset == nullptr) nsub = d;
  B2_CHECK(nsub >= 1 && nsub <= d, B200RBF_EINVAL, "subset size %d out of range", nsub);
  for (int j = 0; subset && j < nsub; ++j)
    B2_CHECK(subset[j] >= 0 && subset[j] < d, B200RBF_EINVAL, "subset index %d out of range", subset[j]);
  for (int k = 0; k < d; ++k) {
    B2_CHECK(lb[k] <= xbest[k] && xbest[k] <= ub[k], B200RBF_EINVAL, "xbest[%d] outside [lb, ub]", k);
    B2_CHECK(kind == 2 || sigma[k] > 0.0, B200RBF_EINVAL, "sigma[%d] must be positive", k);
  }
  B2_CUDA(cudaSetDevice(h->device));
  std::vector<double> ones;
  if (sigma == nullptr) {
    ones.assign(d, 1.0);
    sigma = ones.data();
  }
  return run_generate(h, kind, m, d, xbest, lb, ub, sigma, prob_perturb, subset, nsub, int_mask, seed, cand_dev);
}

int b200rbf_fp64_peaks(b200rbf_handle h, double ms, double* out) {
  B2_CHECK(h && out, B200RBF_EINVAL, "NULL argument");
  B2_CUDA(cudaSetDevice(h->device));
  return run_fp64_peaks(h, ms, out);
}

/* ---- test hooks (not part of the drop-in surface; used by tests/ to localise failures) ---- */
int b200rbf_dbg_lu(b200rbf_handle h, double* A_host, int N, int nrhs, int* ipiv_host, int* info) {
  // A_host: N x (N + nrhs) row-major; factored in place, right-hand sides replaced by L^-1 P b
  B2_CHECK(h && A_host && ipiv_host && info, B200RBF_EINVAL, "NULL argument");
  B2_CUDA(cudaSetDevice(h->device));
  const int rcol0 = (N + 1) & ~1, ncols = rcol0 + nrhs, ld = (ncols + 7) & ~7;
  DevBuf A, piv;
  B2_TRY(A.reserve(((size_t)N * ld + 1024) * sizeof(double)));
  B2_TRY(piv.reserve(((size_t)N + 16) * sizeof(int)));
  B2_CUDA(cudaMemsetAsync(A.p, 0, (size_t)N * ld * sizeof(double), h->stream));
  B2_CUDA(cudaMemsetAsync(piv.p, 0, ((size_t)N + 16) * sizeof(int), h->stream));
  B2_CUDA(cudaMemcpy2DAsync(A.p, (size_t)ld * 8, A_host, (size_t)(N + nrhs) * 8, (size_t)N * 8, N,
                            cudaMemcpyHostToDevice, h->stream));
  if (nrhs > 0)
    B2_CUDA(cudaMemcpy2DAsync(A.as<double>() + rcol0, (size_t)ld * 8, A_host + N, (size_t)(N + nrhs) * 8,
                              (size_t)nrhs * 8, N, cudaMemcpyHostToDevice, h->stream));
  int rc = run_lu_debug(h, A.as<double>(), ld, N, rcol0, ncols, piv.as<int>(), piv.as<int>() + N);
  if (rc == 0) {
    cudaMemcpy2DAsync(A_host, (size_t)(N + nrhs) * 8, A.p, (size_t)ld * 8, (size_t)N * 8, N, cudaMemcpyDeviceToHost,
                      h->stream);
    if (nrhs > 0)
      cudaMemcpy2DAsync(A_host + N, (size_t)(N + nrhs) * 8, A.as<double>() + rcol0, (size_t)ld * 8, (size_t)nrhs * 8, N,
                        cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(ipiv_host, piv.p, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(info, piv.as<int>() + N, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  }
  cudaError_t e = cudaStreamSynchronize(h->stream);
  A.release();
  piv.release();
  if (rc != 0) return rc;
  B2_CUDA(e);
  return 0;
}

int b200rbf_dbg_gemm(b200rbf_handle h, double* C_host, const double* A_host, const double* B_host, int M, int Nc,
                     int K, double* seconds) {
  // C (M x Nc) -= A (M x K) * B (K x Nc), row-major host arrays; K % 16 == 0, Nc even
  B2_CHECK(h && C_host && A_host && B_host, B200RBF_EINVAL, "NULL argument");
  B2_CHECK(K % 16 == 0 && K > 0, B200RBF_EINVAL, "K must be a positive multiple of 16");
  B2_CUDA(cudaSetDevice(h->device));
  const int ldc = (Nc + 1) & ~1, lda = K, ldb = ldc;
  DevBuf C, A, B;
  B2_TRY(C.reserve(((size_t)M * ldc + 256) * 8));
  B2_TRY(A.reserve(((size_t)M * lda + 256) * 8));
  B2_TRY(B.reserve(((size_t)K * ldb + 256) * 8));
  B2_CUDA(cudaMemsetAsync(B.p, 0, ((size_t)K * ldb + 256) * 8, h->stream));
  B2_CUDA(cudaMemcpy2DAsync(C.p, (size_t)ldc * 8, C_host, (size_t)Nc * 8, (size_t)Nc * 8, M, cudaMemcpyHostToDevice,
                            h->stream));
  B2_CUDA(cudaMemcpyAsync(A.p, A_host, (size_t)M * K * 8, cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(cudaMemcpy2DAsync(B.p, (size_t)ldb * 8, B_host, (size_t)Nc * 8, (size_t)Nc * 8, K, cudaMemcpyHostToDevice,
                            h->stream));
  B2_CUDA(cudaEventRecord(h->ev_t0, h->stream));
  int rc = run_gemm_debug(h, C.as<double>(), ldc, A.as<double>(), lda, B.as<double>(), ldb, M, Nc, K);
  cudaEventRecord(h->ev_t1, h->stream);
  if (rc == 0)
    cudaMemcpy2DAsync(C_host, (size_t)Nc * 8, C.p, (size_t)ldc * 8, (size_t)Nc * 8, M, cudaMemcpyDeviceToHost,
                      h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (seconds) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    *seconds = ms * 1e-3;
  }
  C.release();
  A.release();
  B.release();
  if (rc != 0) return rc;
  B2_CUDA(e);
  return 0;
}

}  // extern "C"
