// K1d: RBFInterpolant._fit, incremental branch (rbf.py:132-161) on the device -- row-append extension of the
// null-space Cholesky factorisation of fit_chol.cu -- and the single-launch pipelined triangular solves both use.
//
// The reference extends its pivoted LU of the saddle-point matrix by the new points' rows / columns and a Cholesky
// factorisation of their Schur complement, O(k n^2) instead of O(n^3), falling back to a full refit when that Cholesky
// fails (rbf.py:149-153).  Here the resident factorisation is  M0 = L L^T,  M0 = Pi0 (Phi00 + eta I) Pi0 + sigma Q0 Q0^T,
// P0 = Q0 R0 (fit_chol.cu), of the FIRST n0 points.  A point j added later gets the null-space direction
//     z_j = [-Q0 qt_j ; e_j],   qt_j = R0^-T p_j,        (P_all^T z_j = -R0^T qt_j + p_j = 0)
// which touches only the base points and itself, so with  T = [[Pi0, -Q0 Qt^T], [0, I]]  the matrix
//     M = T^T (Phi + eta I) T + sigma [Q0; 0][Q0; 0]^T
// has M0 as its leading block: its Cholesky factor grows by ONE ROW per point,
//     m   = T^T (Phi + eta I) z_j           = g - Qx (Q0^T g_0),     g = phi(|u_. - u_j|) + eta e_j - Wx qt_j
//     l   = L^-1 m[0:N],   l_NN = sqrt(m_N - l.l)                    (non-positive -> caller refits, like rbf.py:151)
// with Qx = [Q0; Qt] and Wx = [(Phi00 + eta I) Q0 ; Q0^T phi(X0, x_i) rows] kept per point.  The weights are
//     c = T L^-T L^-1 T^T f,   lambda = R0^-1 Q0^T (f_0 - ((Phi + eta I) c)_0)
// where u = L^-1 T^T f also grows by one entry per point.  Z = T restricted to null(P^T) satisfies
// Z^T Z = I + Qt Qt^T with eigenvalues ~ 1 + k / n0, so the conditioning of M stays within that factor of the freshly
// projected matrix; the caller refits from scratch once the point count exceeds 4 n0 (or the pre-allocated capacity).
// Measured against the reference's FRESH fit on the config #2 trace cloud: 6e-12 (tools notes in DESIGN.md).
//
// Triangular solves: one launch, one CTA per 128-row block; CTA b waits (acquire loads on per-block flags) for the
// solution blocks it depends on, applies L_bc x_c as they arrive, multiplies by the stored inverse of its diagonal
// block and publishes x_b.  Logical block numbers are handed out by an atomic ticket, so a CTA only ever waits for
// CTAs that were scheduled before it: no co-residency assumption, no deadlock.  Critical path ~1.5 us per block
// instead of two launches (~8-13 us) per block in the first version.
#include "common.cuh"

namespace b200rbf {

namespace {

constexpr int kTB = 128;            // block size of the solves == kNB of fit_chol.cu (diagonal-block inverses)
constexpr int kTrsvThreads = 512;
constexpr int kLs = kTB + 1;        // padded stride of the staged inverse block
constexpr unsigned kSpinLimit = 1u << 22;

#define LAUNCHED(h)               \
  do {                            \
    (h)->launches++;              \
    B2_CUDA(cudaGetLastError());  \
  } while (0)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// x = L^-1 b (TRANS == false) or L^-T b (TRANS == true) for the leading N x N lower-triangular L (row-major, ld),
// Linv = inverses of its kTB x kTB diagonal blocks.  grid = ceil(N / kTB) CTAs.  b and x must not alias.
template <bool TRANS>
__global__ void __launch_bounds__(kTrsvThreads, 1) trsv_pipe_kernel(const double* __restrict__ L, int ld,
                                                                    const double* __restrict__ Linv, int N,
                                                                    const double* __restrict__ b, double* x,
                                                                    TrsvSync* sy, unsigned epoch) {
  extern __shared__ double sm[];
  double* sL = sm;                      // kTB x kLs : inverse of the diagonal block
  double* part = sL + kTB * kLs;        // 16 x kTB  : cross-warp partial sums
  double* r = part + 16 * kTB;          // kTB       : running right-hand side of this block
  double* xc = r + kTB;                 // kTB       : the solution block being applied
  __shared__ int s_seq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x;
  if (tid == 0) s_seq = (int)atomicAdd(&sy->ticket, 1u);
  __syncthreads();
  const int seq = s_seq;
  const int bi = TRANS ? nblk - 1 - seq : seq;
  const int r0 = bi * kTB;
  const int rows = min(kTB, N - r0);
  {
    const double* Lb = Linv + (size_t)bi * kTB * kTB;
    for (int e = tid; e < kTB * kTB; e += kTrsvThreads) {
      const int i = e >> 7, c = e & (kTB - 1);
      sL[i * kLs + c] = (i < rows && c < rows) ? Lb[e] : 0.0;
    }
    if (tid < kTB) r[tid] = tid < rows ? b[r0 + tid] : 0.0;
  }
  __syncthreads();
  for (int s = 0; s < seq; ++s) {
    const int c = TRANS ? nblk - 1 - s : s;
    const int c0 = c * kTB;
    const int crows = min(kTB, N - c0);
    // the L block this step needs, fetched BEFORE waiting for x_c: forward rows r0 + i, columns c0 + j;
    // transpose rows c0 + i, columns r0 + j.  warp = 8 rows i, lane = columns j = lane + 32 q (coalesced).
    double a[8][4];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int i = warp * 8 + ii;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        const bool ok = TRANS ? (i < crows && j < rows) : (i < rows && j < crows);
        const size_t off = TRANS ? (size_t)(c0 + i) * ld + r0 + j : (size_t)(r0 + i) * ld + c0 + j;
        a[ii][q] = ok ? L[off] : 0.0;
      }
    }
    if (tid == 0) {
      unsigned spins = 0;
      while (ld_acquire_u32(&sy->flag[c]) != epoch) {
        if (++spins > kSpinLimit) {
          atomicExch(&sy->err, 1u);
          break;
        }
        if (spins > 64) __nanosleep(100);
      }
    }
    __syncthreads();
    if (tid < kTB) xc[tid] = tid < crows ? __ldcg(x + c0 + tid) : 0.0;
    __syncthreads();
    if (!TRANS) {
      double xv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) xv[q] = xc[lane + 32 * q];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        double acc = fma(a[ii][0], xv[0], a[ii][1] * xv[1]) + fma(a[ii][2], xv[2], a[ii][3] * xv[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) r[warp * 8 + ii] -= acc;
      }
    } else {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const double yv = xc[warp * 8 + ii];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fma(a[ii][q], yv, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) part[warp * kTB + lane + 32 * q] = acc[q];
      __syncthreads();
      if (tid < kTB) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 16; ++w) t += part[w * kTB + tid];
        r[tid] -= t;
      }
    }
  }
  __syncthreads();
  {
    const int o = tid & (kTB - 1), p = tid >> 7;  // 4 partial sums per output
    double a0 = 0.0;
    if (o < rows) {
      if (!TRANS) {
        for (int k = p; k <= o; k += 4) a0 = fma(sL[o * kLs + k], r[k], a0);
      } else {
        for (int i = o + p; i < rows; i += 4) a0 = fma(sL[i * kLs + o], r[i], a0);
      }
    }
    part[p * kTB + o] = a0;
  }
  __syncthreads();
  if (tid < rows) __stcg(x + r0 + tid, (part[tid] + part[kTB + tid]) + (part[2 * kTB + tid] + part[3 * kTB + tid]));
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    st_release_u32(&sy->flag[bi], epoch);
    const unsigned old = atomicAdd(&sy->done, 1u);
    if (old == (unsigned)nblk - 1u) {  // last CTA out: every ticket has been taken, re-arm for the next launch
      sy->ticket = 0u;
      sy->done = 0u;
    }
  }
}

constexpr size_t kTrsvSmem = ((size_t)kTB * kLs + 16 * kTB + 2 * kTB) * sizeof(double);

// ---------------------------------------------------------------------------------------------------- extension
__device__ __forceinline__ double ext_phi_r2(int kern, double r2) {  // same arithmetic as fit_kernels.cu's assembly
  if (kern == B200RBF_KERNEL_CUBIC) return r2 * sqrt(r2);
  if (kern == B200RBF_KERNEL_TPS) {
    const double eps2 = 4.930380657631323783823303533017413935457e-32;
    const double c = fmax(r2, eps2);
    return 0.5 * c * log(c);
  }
  return sqrt(r2);
}

// out[a] = sum_{i < n} Q[i][a] v[i]  for a < tp, whole CTA; red: blockDim doubles of shared scratch.  Ends with a barrier.
__device__ void cta_qt_vec(const double* __restrict__ Q, int n, int tp, const double* v, double* out, double* red) {
  const int tid = threadIdx.x, a = tid % tp, groups = blockDim.x / tp, g = tid / tp;
  double s = 0.0;
  if (g < groups)
    for (int i = g; i < n; i += groups) s = fma(Q[(size_t)i * tp + a], v[i], s);
  red[tid] = (g < groups) ? s : 0.0;
  __syncthreads();
  if (tid < tp) {
    double t = 0.0;
    for (int q = 0; q < groups; ++q) t += red[q * tp + tid];
    out[tid] = t;
  }
  __syncthreads();
}

// Everything that precedes the triangular solve for new point j = N (its coordinates are already in Xu row N), in two
// row-parallel launches (the first version was one 1024-thread CTA walking all rows five times: 60 us at N = 1000):
//   A:  qt = Rinv^T p_j -> Qx[N];  col_i = phi(|u_i - u_j|) (+ eta at i = N);  g_i = col_i - Wx[i].qt (i < N);
//       per-CTA partial sums over the base rows of  Q0^T col  (-> w = Wx[N])  and  Q0^T g  (-> s)
//   B:  w, s from the partials (fixed order: deterministic);  g_N = col_N - w.qt;  mrow[i] = g_i - Qx[i].s (i <= N);
//       Wx[N] = w;  bvec[N] = f_N - qt.fq
constexpr int kExtRows = 256;     // rows (= threads) per CTA of the row-parallel passes
constexpr int kExtThreads = 1024;  // the single-CTA reductions
__global__ void __launch_bounds__(kExtRows) ext_row_a_kernel(const double* __restrict__ Xu, int d, int n0, int N, int kern,
                                                             double eta, const double* __restrict__ Rinv, int t, int tp,
                                                             double* Qx, const double* __restrict__ Wx, double* gbuf,
                                                             double* __restrict__ part, int exact) {
  extern __shared__ double sh[];
  double* uj = sh;             // d
  double* p = uj + d;          // tp
  double* qt = p + tp;         // tp
  double* scol = qt + tp;      // kExtRows
  double* sg = scol + kExtRows;  // kExtRows
  const int tid = threadIdx.x;
  for (int k = tid; k < d; k += kExtRows) uj[k] = Xu[(size_t)N * d + k];
  __syncthreads();
  for (int a = tid; a < tp; a += kExtRows) p[a] = a == 0 ? 1.0 : (a < t ? uj[a - 1] : 0.0);
  __syncthreads();
  for (int a = tid; a < tp; a += kExtRows) {  // qt_a = sum_{k <= a} Rinv[k][a] p_k   (R0^T qt = p)
    double s = 0.0;
    if (a < t)
      for (int k = 0; k <= a; ++k) s = fma(Rinv[k * tp + a], p[k], s);
    qt[a] = s;
  }
  __syncthreads();
  const int i = blockIdx.x * kExtRows + tid;
  double col = 0.0, g = 0.0;
  if (i <= N) {
    const double* ui = Xu + (size_t)i * d;
    double r2 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = ui[k] - uj[k];
      r2 = exact ? __dadd_rn(r2, __dmul_rn(df, df)) : fma(df, df, r2);
    }
    col = ext_phi_r2(kern, r2);
    if (i == N) {
      col += eta;
      g = col;  // w.qt is subtracted in pass B
      for (int a = 0; a < tp; ++a) Qx[(size_t)N * tp + a] = qt[a];
    } else {
      const double* wr = Wx + (size_t)i * tp;
      double s = 0.0;
      for (int a = 0; a < t; ++a) s = fma(wr[a], qt[a], s);
      g = col - s;
    }
    gbuf[i] = g;
  }
  const bool base = i < n0;
  scol[tid] = base ? col : 0.0;
  sg[tid] = base ? g : 0.0;
  __syncthreads();
  // per-CTA partial sums over this CTA's base rows, thread = (row group, column): 8 groups x 32 columns walk the rows
  // 8 apart (coalesced across the columns of a Q row), then the groups are added in a fixed order.  (The first version
  // gave each of the tp <= 32 columns one thread walking all 256 rows: ~100 us of dependent L2 loads per extension,
  // the largest single item of the 0.3 ms at n = 2700.)
  const int r0 = blockIdx.x * kExtRows;
  if (r0 < n0) {  // block-uniform
    const int rows = min(kExtRows, n0 - r0);
    double* red = sg + kExtRows;  // 2 x 8 x 32
    const int cl = tid & 31, rg = tid >> 5;
    for (int a0 = 0; a0 < tp; a0 += 32) {
      const int a = a0 + cl;
      double sw = 0.0, ss = 0.0;
      if (a < tp)
        for (int r = rg; r < rows; r += kExtRows / 32) {
          const double q = Qx[(size_t)(r0 + r) * tp + a];
          sw = fma(q, scol[r], sw);
          ss = fma(q, sg[r], ss);
        }
      red[rg * 32 + cl] = sw;
      red[kExtRows + rg * 32 + cl] = ss;
      __syncthreads();
      if (tid < 32 && a < tp) {
        double w = 0.0, s2 = 0.0;
        for (int g = 0; g < kExtRows / 32; ++g) {
          w += red[g * 32 + tid];
          s2 += red[kExtRows + g * 32 + tid];
        }
        part[((size_t)blockIdx.x * 2 + 0) * tp + a] = w;
        part[((size_t)blockIdx.x * 2 + 1) * tp + a] = s2;
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kExtRows) ext_row_b_kernel(int n0, int N, int t, int tp, const double* Qx, double* Wx,
                                                             const double* __restrict__ part,
                                                             const double* __restrict__ fq, const double* __restrict__ f,
                                                             const double* gbuf, double* __restrict__ mrow,
                                                             double* __restrict__ bvec) {
  extern __shared__ double sh[];
  double* w = sh;        // tp
  double* sv = w + tp;   // tp
  double* qt = sv + tp;  // tp
  __shared__ double s_wq;
  const int tid = threadIdx.x;
  const int nbase = (n0 + kExtRows - 1) / kExtRows;  // CTAs of pass A that held base rows
  for (int a = tid; a < tp; a += kExtRows) {
    double sw = 0.0, ss = 0.0;
    for (int b = 0; b < nbase; ++b) {
      sw += part[((size_t)b * 2 + 0) * tp + a];
      ss += part[((size_t)b * 2 + 1) * tp + a];
    }
    w[a] = sw;
    sv[a] = ss;
    qt[a] = Qx[(size_t)N * tp + a];
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int a = 0; a < t; ++a) s = fma(w[a], qt[a], s);
    s_wq = s;
  }
  __syncthreads();
  const int i = blockIdx.x * kExtRows + tid;
  if (i <= N) {
    const double* qr = Qx + (size_t)i * tp;
    double s = 0.0;
    for (int a = 0; a < t; ++a) s = fma(qr[a], sv[a], s);
    const double g = gbuf[i] - (i == N ? s_wq : 0.0);
    mrow[i] = g - s;
  }
  if (blockIdx.x == (unsigned)(N / kExtRows)) {  // the CTA that owns row N
    for (int a = tid; a < tp; a += kExtRows) Wx[(size_t)N * tp + a] = w[a];
    if (tid == 0) {
      double s = 0.0;
      for (int a = 0; a < t; ++a) s = fma(qt[a], fq[a], s);
      bvec[N] = f[N] - s;
    }
  }
}

// After l = L^-1 m[0:N] landed in row N of A: the pivot, u_N, and the new row of the diagonal-block inverse.
__global__ void __launch_bounds__(512) finish_row_kernel(double* __restrict__ A, int ld, int N,
                                                         const double* __restrict__ mrow,
                                                         const double* __restrict__ bvec, double* __restrict__ u,
                                                         double* __restrict__ Linv, int* __restrict__ info) {
  __shared__ double r1[512], r2[512];
  __shared__ double s_inv;
  const int tid = threadIdx.x;
  const double* l = A + (size_t)N * ld;
  double a = 0.0, b = 0.0;
  for (int i = tid; i < N; i += 512) {
    const double v = l[i];
    a = fma(v, v, a);
    b = fma(v, u[i], b);
  }
  r1[tid] = a;
  r2[tid] = b;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (tid < o) {
      r1[tid] += r1[tid + o];
      r2[tid] += r2[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double md = mrow[N];
    const double dd = md - r1[0];
    // a new point that is numerically dependent on the old ones: let the caller refit (rbf.py:151-153)
    const bool ok = dd > 0.0 && dd > 1e-13 * md && dd < 1.7e308;
    if (!ok && *info == 0) *info = N + 1;
    const double lnn = ok ? sqrt(dd) : 1.0;
    A[(size_t)N * ld + N] = lnn;
    u[N] = (bvec[N] - r2[0]) / lnn;
    s_inv = 1.0 / lnn;
  }
  __syncthreads();
  const int b0 = (N / kTB) * kTB, rr = N - b0;
  double* Lb = Linv + (size_t)(N / kTB) * kTB * kTB;
  if (rr == 0)
    for (int e = tid; e < kTB * kTB; e += 512) Lb[e] = 0.0;  // a new diagonal block starts
  __syncthreads();
  const double inv = s_inv;
  if (tid < kTB) {
    double v = 0.0;
    if (tid < rr) {
      double s = 0.0;
      for (int k = tid; k < rr; ++k) s = fma(l[b0 + k], Lb[k * kTB + tid], s);
      v = -inv * s;
    } else if (tid == rr) {
      v = inv;
    }
    Lb[rr * kTB + tid] = v;
  }
}

// Rinv = R^-1 for the upper-triangular t x t R (leading dimension tp); thread = column, back substitution
__global__ void rinv_kernel(const double* __restrict__ R, int tp, int t, double* __restrict__ Rinv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= tp) return;
  for (int i = tp - 1; i >= 0; --i) {
    double v = 0.0;
    if (j < t && i <= j) {
      v = (i == j) ? 1.0 : 0.0;
      for (int k = i + 1; k <= j; ++k) v = fma(-R[i * tp + k], Rinv[k * tp + j], v);
      v /= R[i * tp + i];
    }
    Rinv[i * tp + j] = v;
  }
}

// c_out[t + i] = y[i] - (i < n0 ? Q0[i].v : 0),  v = Qx^T y  (N rows), c_out[0 .. t) = 0;  one CTA computes v, then the
// elementwise pass
__global__ void __launch_bounds__(kExtThreads) ext_weights_kernel(const double* __restrict__ Qx, int n0, int N, int t,
                                                                  int tp, const double* __restrict__ y,
                                                                  double* __restrict__ c_out) {
  extern __shared__ double sh[];
  double* red = sh;
  double* v = red + kExtThreads;
  double* c = c_out + t;
  cta_qt_vec(Qx, N, tp, y, v, red);
  if (threadIdx.x < t) c_out[threadIdx.x] = 0.0;
  for (int i = threadIdx.x; i < N; i += kExtThreads) {
    double s = 0.0;
    if (i < n0) {
      const double* qr = Qx + (size_t)i * tp;
      for (int a = 0; a < t; ++a) s = fma(qr[a], v[a], s);
    }
    c[i] = y[i] - s;
  }
}

// b = T^T f over all N points: fq = Q0^T f_0, b[i] = f[i] - Qx[i].fq
__global__ void __launch_bounds__(kExtThreads) ext_rhs_kernel(const double* __restrict__ Qx, int n0, int N, int t, int tp,
                                                              const double* __restrict__ f, double* __restrict__ fq,
                                                              double* __restrict__ bvec) {
  extern __shared__ double sh[];
  double* red = sh;
  double* v = red + kExtThreads;
  cta_qt_vec(Qx, n0, tp, f, v, red);
  if (threadIdx.x < tp) fq[threadIdx.x] = v[threadIdx.x];
  for (int i = threadIdx.x; i < N; i += kExtThreads) {
    const double* qr = Qx + (size_t)i * tp;
    double s = 0.0;
    for (int a = 0; a < t; ++a) s = fma(qr[a], v[a], s);
    bvec[i] = f[i] - s;
  }
}

}  // namespace

// the flags of the pipelined solves; its `info` word doubles as the factorisation status of fit_chol.cu / this file
int ensure_fit_sync(b200rbf_ctx* h) {
  FitState& F = h->fit;
  if (F.sync.p == nullptr) {
    B2_TRY(F.sync.reserve(sizeof(TrsvSync)));
    B2_CUDA(cudaMemsetAsync(F.sync.p, 0, sizeof(TrsvSync), h->stream));
  }
  return 0;
}

int trsv_pipelined(b200rbf_ctx* h, const double* L, int ld, const double* Linv, int N, const double* b, double* x,
                   bool transpose) {
  FitState& F = h->fit;
  B2_TRY(ensure_fit_sync(h));
  const int nblk = (N + kTB - 1) / kTB;
  B2_CHECK(nblk >= 1 && nblk <= kTrsvMaxBlocks, B200RBF_EINVAL, "triangular solve: %d rows out of range", N);
  static DeviceOnce attr;
  if (attr.pending(h->device)) {
    B2_CUDA(cudaFuncSetAttribute(trsv_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsvSmem));
    B2_CUDA(cudaFuncSetAttribute(trsv_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsvSmem));
    attr.done(h->device);
  }
  unsigned epoch = ++F.trsv_epoch;
  if (epoch == 0u) epoch = ++F.trsv_epoch;  // 0 is the cleared state of the flags
  if (transpose)
    trsv_pipe_kernel<true><<<nblk, kTrsvThreads, kTrsvSmem, h->stream>>>(L, ld, Linv, N, b, x, F.sync.as<TrsvSync>(), epoch);
  else
    trsv_pipe_kernel<false><<<nblk, kTrsvThreads, kTrsvSmem, h->stream>>>(L, ld, Linv, N, b, x, F.sync.as<TrsvSync>(), epoch);
  LAUNCHED(h);
  return 0;
}

int ext_make_rinv(b200rbf_ctx* h, const CholWork& w) {
  rinv_kernel<<<(w.tp + 63) / 64, 64, 0, h->stream>>>(w.R, w.tp, w.t, w.Rinv);
  LAUNCHED(h);
  return 0;
}

static size_t ext_smem(int tp, int d) { return ((size_t)kExtThreads + 4 * (size_t)tp + (size_t)d + 8) * sizeof(double); }

// b = T^T f for all resident points, u = L^-1 b  (right-hand side changed, or first use after the base fit)
int ext_refresh_rhs(b200rbf_ctx* h, const double* f_dev) {
  ExtState& E = h->fit.ext;
  const CholWork& w = E.w;
  ext_rhs_kernel<<<1, kExtThreads, ext_smem(w.tp, 0), h->stream>>>(w.P, E.n0, E.N, w.t, w.tp, f_dev, w.fq, w.bvec);
  LAUNCHED(h);
  B2_TRY(trsv_pipelined(h, h->fit.A.as<double>(), w.ld, w.Linv, E.N, w.bvec, w.u, false));
  return 0;
}

// append points [E.N, n) (their unit-box rows are already in E.Xu, values f_dev) to the resident factorisation and solve for the
// RBF weights of all n points -> c_out = [0_t; weights].  info_dev != 0 afterwards: a pivot failed, the state is invalid.
int chol_extend_weights(b200rbf_ctx* h, const double* f_dev, int n, bool rhs_changed, int* info_dev, double* c_out) {
  ExtState& E = h->fit.ext;
  const CholWork& w = E.w;
  double* A = h->fit.A.as<double>();
  const int d = E.d;
  if (rhs_changed) B2_TRY(ext_refresh_rhs(h, f_dev));
  for (int j = E.N; j < n; ++j) {
    const int ctas = (j + 1 + kExtRows - 1) / kExtRows;
    const size_t sm_a = ((size_t)d + 2 * (size_t)w.tp + 4 * kExtRows) * sizeof(double), sm_b = 3 * (size_t)w.tp * sizeof(double);
    ext_row_a_kernel<<<ctas, kExtRows, sm_a, h->stream>>>(E.Xu.as<double>(), d, E.n0, j, E.kernel, E.eta, w.Rinv, w.t, w.tp, w.P,
                                                          w.W, w.gbuf, w.part, h->exact_dist ? 1 : 0);
    LAUNCHED(h);
    ext_row_b_kernel<<<ctas, kExtRows, sm_b, h->stream>>>(E.n0, j, w.t, w.tp, w.P, w.W, w.part, w.fq, f_dev, w.gbuf, w.mrow,
                                                          w.bvec);
    LAUNCHED(h);
    B2_TRY(trsv_pipelined(h, A, w.ld, w.Linv, j, w.mrow, A + (size_t)j * w.ld, false));
    finish_row_kernel<<<1, 512, 0, h->stream>>>(A, w.ld, j, w.mrow, w.bvec, w.u, w.Linv, info_dev);
    LAUNCHED(h);
  }
  E.N = n;
  B2_TRY(trsv_pipelined(h, A, w.ld, w.Linv, n, w.u, w.v1, true));
  ext_weights_kernel<<<1, kExtThreads, ext_smem(w.tp, 0), h->stream>>>(w.P, E.n0, n, w.t, w.tp, w.v1, c_out);
  LAUNCHED(h);
  return 0;
}

}  // namespace b200rbf
