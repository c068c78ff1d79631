// K1 (fast path): null-space Cholesky solve of the RBF saddle-point system, for the conditionally positive
// definite kernels (cubic, thin-plate spline) with the linear tail.
//
// The reference solves  [[0, P^T], [P, Phi + eta I]] [lambda; c] = [0; f]  by pivoted LU (rbf.py:110-130, 164-168).
// Because P^T c = 0 and Phi is positive definite on null(P^T), the same solution is
//     c      = Pi M^-1 Pi f,        M = Pi (Phi + eta I) Pi + sigma Q Q^T,   Pi = I - Q Q^T,   P = Q R  (thin QR)
//     lambda = R^-1 Q^T (f - (Phi + eta I) c)
// where M is symmetric positive definite with the spectrum of the projected matrix (plus sigma), so it has a
// Cholesky factorisation: no pivot search, no row interchanges, half the flops of the LU (N^3/3), and the trailing
// updates -- still the one real dense contraction -- run on the FP64 tensor cores through the same DMMA GEMM,
// restricted to the lower triangle.  Any failure (P rank deficient, M not numerically positive definite) is reported
// so that the caller falls back to the pivoted LU, mirroring the reference's own Cholesky -> LU fallback
// (rbf.py:149-153).  Both solvers are backward stable, so predictions agree to the reference's reorder noise.
//
// Steps (all on the handle's stream, all row-major):
//   1. Phi + eta I -> A (n16 x ld, n16 = n rounded to 32)                                    assemble_plain
//   2. P = [1, Xu] (n x tp, tp = t rounded to 16), Q, R by CholeskyQR2                       gram / chol_small / apply_rinv
//   3. W = (Phi + eta I) Q                                                                   DMMA GEMM (K = n16)
//   4. S = Q^T W,  M = A - [Q | W | -Q (S + sigma I)] [W^T ; Q^T ; Q^T]  (lower triangle)    DMMA GEMM (K = 3 tp)
//   5. two-level blocked Cholesky: 128-column steps (diagonal block factored AND inverted in one CTA's shared memory,
//      the panel below it by one DMMA GEMM with the inverse) inside 512-column outer panels whose trailing update is
//      one lower-triangle DMMA GEMM with K = 512
//   6. c = Pi L^-T L^-1 Pi f   (strip GEMV + 128 x 128 matrix-vector products with the diagonal-block inverses)
//   7. lambda = R^-1 (Q^T f - W^T c): the projected residual needs W from step 3 only, no second pass over Phi
#include "common.cuh"

namespace b200rbf {

int gemm_sub_ex(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int Nc,
                int K, int lower);
int assemble_plain(b200rbf_ctx* h, const double* Xu_dev, int n, int d, int kernel, double eta, double* A, int ld,
                   int ncols);
int strip_gemv(b200rbf_ctx* h, const double* A, int ld, int r0, int rows, int c0, int c1, const double* x, double* y);
int trsv_pipelined(b200rbf_ctx* h, const double* L, int ld, const double* Linv, int N, const double* b, double* x,
                   bool transpose);
int ext_make_rinv(b200rbf_ctx* h, const CholWork& w);

namespace {

constexpr int kNB = 128;     // columns factored per potrf step
constexpr int kOuter = 512;  // columns whose trailing update is applied at once (K of the big DMMA GEMM)

#define LAUNCHED(h)               \
  do {                            \
    (h)->launches++;              \
    B2_CUDA(cudaGetLastError());  \
  } while (0)

// P (rows_pad x tp): [1, u_1 .. u_d, 0 ..]; rows >= n are zero
__global__ void build_p_kernel(const double* __restrict__ Xu, int n, int d, int t, double* __restrict__ P, int tp,
                               int rows_pad) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)rows_pad * tp) return;
  const int i = (int)(e / tp), c = (int)(e - (long long)i * tp);
  double v = 0.0;
  if (i < n && c < t) v = (c == 0) ? 1.0 : Xu[(size_t)i * d + c - 1];
  P[e] = v;
}

// partial[b] = X[rows of slab b]^T Y[rows of slab b]   (tp x tp each), slabs of 64 rows
__global__ void __launch_bounds__(256) gram_partial_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                           int n, int tp, double* __restrict__ partial) {
  extern __shared__ double sm[];
  double* sx = sm;
  double* sy = sm + 64 * tp;
  const int r0 = blockIdx.x * 64;
  const int rows = min(64, n - r0);
  for (int e = threadIdx.x; e < 64 * tp; e += blockDim.x) {
    const int r = e / tp;
    sx[e] = r < rows ? X[(size_t)(r0 + r) * tp + (e - r * tp)] : 0.0;
    sy[e] = r < rows ? Y[(size_t)(r0 + r) * tp + (e - r * tp)] : 0.0;
  }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * tp * tp;
  for (int e = threadIdx.x; e < tp * tp; e += blockDim.x) {
    const int a = e / tp, b = e - a * tp;
    double s = 0.0;
    for (int r = 0; r < 64; ++r) s = fma(sx[r * tp + a], sy[r * tp + b], s);
    out[e] = s;
  }
}
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nparts, int tp, double* __restrict__ G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= tp * tp) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * tp * tp + e];  // fixed order: deterministic
  G[e] = s;
}

// upper Cholesky factor of the leading t x t block of G (G = R^T R); pad rows/cols of R are the identity
__global__ void __launch_bounds__(256) chol_small_kernel(const double* __restrict__ G, int tp, int t,
                                                         double* __restrict__ R, int* __restrict__ info) {
  extern __shared__ double s[];  // tp x tp working copy
  for (int e = threadIdx.x; e < tp * tp; e += blockDim.x) s[e] = G[e];
  __syncthreads();
  for (int k = 0; k < t; ++k) {
    if (threadIdx.x == 0) {
      const double dgl = s[k * tp + k];
      if (!(dgl > 0.0) && *info == 0) *info = -(k + 1);  // negative: failure in the small QR factor
      s[k * tp + k] = sqrt(dgl);
    }
    __syncthreads();
    const double rkk = s[k * tp + k];
    for (int j = k + 1 + threadIdx.x; j < t; j += blockDim.x) s[k * tp + j] /= rkk;
    __syncthreads();
    const int w = t - k - 1;
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
      const int i = k + 1 + e / w, j = k + 1 + e % w;
      if (j >= i) s[i * tp + j] = fma(-s[k * tp + i], s[k * tp + j], s[i * tp + j]);
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < tp * tp; e += blockDim.x) {
    const int i = e / tp, j = e - i * tp;
    double v = 0.0;
    if (i < t && j < t) v = j >= i ? s[e] : 0.0;
    else if (i == j) v = 1.0;
    R[e] = v;
  }
}

// X <- X R^-1 row by row (solve q R = x with R upper triangular); rows live in global / L1
__global__ void apply_rinv_kernel(double* __restrict__ X, int n, int tp, int t, const double* __restrict__ R) {
  extern __shared__ double sr[];
  for (int e = threadIdx.x; e < tp * tp; e += blockDim.x) sr[e] = R[e];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* x = X + (size_t)i * tp;
  for (int j = 0; j < t; ++j) {
    double v = x[j];
    for (int k = 0; k < j; ++k) v = fma(-x[k], sr[k * tp + j], v);
    x[j] = v / sr[j * tp + j];
  }
}

// C = A B for tp x tp matrices
__global__ void small_mm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                int tp) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= tp * tp) return;
  const int i = e / tp, j = e - i * tp;
  double s = 0.0;
  for (int k = 0; k < tp; ++k) s = fma(A[i * tp + k], B[k * tp + j], s);
  C[e] = s;
}

// Acat (n x 3 tp) = [Q | W | -(Q T)],  T = S + sigma I on the leading t x t block
__global__ void build_acat_kernel(const double* __restrict__ Q, const double* __restrict__ W,
                                  const double* __restrict__ S, int n, int tp, int t, double sigma,
                                  double* __restrict__ Acat) {
  extern __shared__ double st[];
  for (int e = threadIdx.x; e < tp * tp; e += blockDim.x) {
    const int i = e / tp, j = e - i * tp;
    st[e] = (i < t && j < t) ? S[e] + (i == j ? sigma : 0.0) : 0.0;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* q = Q + (size_t)i * tp;
  const double* w = W + (size_t)i * tp;
  double* out = Acat + (size_t)i * 3 * tp;
  for (int j = 0; j < tp; ++j) {
    out[j] = q[j];
    out[tp + j] = w[j];
    double s = 0.0;
    for (int k = 0; k < t; ++k) s = fma(q[k], st[k * tp + j], s);
    out[2 * tp + j] = -s;
  }
}

// dst[c][r] = src[r][c] for r < rows, c < cols (32 x 32 tiles through shared memory)
__global__ void transpose_kernel(const double* __restrict__ src, int rows, int cols, int lds, double* __restrict__ dst,
                                 int ldd) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int r = r0 + dy, c = c0 + threadIdx.x;
    tile[dy][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * lds + c] : 0.0;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int c = c0 + dy, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(size_t)c * ldd + r] = tile[threadIdx.x][dy];
  }
}

// rows [n, n16) of the n16 x n16 matrix become rows of the identity (they decouple from the factorisation)
__global__ void pad_identity_kernel(double* __restrict__ A, int ld, int n, int n16) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = n16 - n;
  if (e >= rows * n16) return;
  const int i = n + e / n16, j = e % n16;
  A[(size_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
}

// 1 / x for the pivots: hardware seed (MUFU.RCP64H, ~2^-20) + two Newton steps = 4 dependent FMAs.  The IEEE division
// is ~10 dependent FP64 operations, and this value heads the critical path of every column of the factorisation.
// A non-positive or non-finite pivot yields a value that fails the `> 0` test downstream (inf, negative or NaN).
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(fma(-x, y, 1.0), y, y);
  y = fma(fma(-x, y, 1.0), y, y);
  return y;
}

// Cholesky of the nb x nb diagonal block at (j0, j0) (nb a multiple of 32) AND the inverse of its factor, one CTA.
// Factorisation: right-looking in 32-column panels held in registers, ONE barrier per column, rank-32 register-tiled
// updates between panels (see the loop).  Inverse by blocks of 32: the diagonal sub-blocks by forward substitution in
// registers (warp per sub-block, lane = column), the sub-blocks at block distance 1, 2, 3 from
// X_IJ = -X_II sum_K L_IK X_KJ with 4 x 4 register tiles.  Writes L back into A and L^-1 / L^-T (kNB x kNB, zero
// padded) for the panel GEMM and the triangular solves.  Phase times at nb = 128 (tools/potrf_probe.cu): load 8 us,
// factorisation 36 us, inverse 24 us, stores 8 us; the first version (three barriers per column, whole-block
// rank-1 sweeps through shared memory) took 125 us for the factorisation alone.
constexpr int kPiThreads = 512;
__global__ void __launch_bounds__(kPiThreads) potrf_inv_kernel(double* __restrict__ A, int ld, int j0, int nb,
                                                               double* __restrict__ Linv, double* __restrict__ LinvT,
                                                               int* __restrict__ info) {
  // S (kNB x (kNB + 1)): L in the lower triangle (diagonal included); X = L^-1 strictly below its diagonal is kept
  // TRANSPOSED in the strict upper triangle (X[i][j] at S[j][i], i > j) and its diagonal is rD.  Tb: three 32 x 33
  // scratch blocks for the partial sums T_IJ.
  extern __shared__ double sm[];
  __shared__ double rD[kNB];  // reciprocals of the diagonal of L = diagonal of L^-1
  constexpr int st = kNB + 1;
  double* S = sm;
  double* Tb = sm + kNB * st;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 8
  for (int e = tid; e < kNB * kNB; e += kPiThreads) {
    const int i = e >> 7, c = e & (kNB - 1);
    S[i * st + c] = (i < nb && c <= i) ? A[(size_t)(j0 + i) * ld + j0 + c] : 0.0;
  }
  __syncthreads();
#if defined(PROBE_STOP) && PROBE_STOP == 4
  return;
#endif
  const int nsub = nb / 32;
  // ---- factorisation: right-looking in panels of 32 columns.  thread = (row i, column group q): for the duration of a
  // panel it keeps its 8 entries S[i][c0 + 8 q .. + 7] in REGISTERS.  Per column k the owners of that column publish
  // it (still unscaled) through a double-buffered shared-memory vector -- ONE barrier per column -- and every thread
  // applies  S[i][j] -= S[i][k] S[j][k] / pivot_k  to its registers: per column a warp issues ~10 broadcast reads
  // instead of the ~48 shared-memory wavefronts of a read-modify-write sweep (measured: the sweep made this phase
  // shared-memory bound, ~1000 cycles per column against a ~250-cycle dependency chain).  After the panel the rest of
  // the block gets its rank-32 update from 4 x 4 register tiles (rows / columns interleaved so that lanes walk
  // consecutive rows: conflict-free; 8 shared-memory reads per 16 FMAs).  The scaling by 1 / sqrt(pivot_k) happens
  // once at the end; meanwhile rD holds the reciprocal pivots.
  {
    __shared__ double colbuf[2][kNB];
    const int i = tid & (kNB - 1), q = tid >> 7;
    bool bad = false;
    for (int c0 = 0; c0 < nb; c0 += 32) {
      const int cend = c0 + 32;
      double reg[8];
      double* mine = S + i * st + c0 + 8 * q;
#pragma unroll
      for (int c = 0; c < 8; ++c) reg[c] = mine[c];
      for (int k8 = 0; k8 < 4; ++k8) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = c0 + 8 * k8 + kk;
          double* cb = colbuf[kk & 1];
          if (q == k8 && i >= k) cb[i] = reg[kk];
          __syncthreads();  // also orders the reads of step k - 1 before the writes of step k + 1 into the same buffer
          const double piv = cb[k];
          bad |= !(piv > 0.0);
          const double rp = fast_rcp(piv);
          if (i == k && q == k8) rD[k] = rp;
          if (i > k && i < nb && q >= k8) {
            const double t = cb[i] * rp;
            const double* cj = cb + c0 + 8 * q;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (q > k8 || c > kk) reg[c] = fma(-t, cj[c], reg[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c0 + 8 * q + c <= i) mine[c] = reg[c];
      __syncthreads();
      // rank-32 update of the m x m block right of / below the panel: i = ti + (m/4) u, j = tj + (m/4) v
      const int m = nb - cend;
      if (m > 0) {
        const int q4 = m / 4;
        for (int t = tid; t < q4 * q4; t += kPiThreads) {
          const int ti = t % q4, tj = t / q4;
          const double* pi = S + (cend + ti) * st + c0;
          const double* pj = S + (cend + tj) * st + c0;
          double acc[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll 4
          for (int k = 0; k < 32; ++k) {
            const double rp = rD[c0 + k];
            double ai[4], bj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              ai[u] = pi[u * q4 * st + k] * rp;
              bj[u] = pj[u * q4 * st + k];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < 4; ++v) acc[u][v] = fma(ai[u], bj[v], acc[u][v]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int ii = ti + q4 * u, jj = tj + q4 * v;
              if (jj <= ii) S[(cend + ii) * st + cend + jj] -= acc[u][v];
            }
        }
        __syncthreads();
      }
    }
    // scale: L[i][k] = S[i][k] / sqrt(pivot_k), L[k][k] = sqrt(pivot_k)
    if (tid < nb) {
      const double dg = S[tid * st + tid];
      if (!(dg > 0.0) || dg > 1.7e308) bad = true;  // zero, negative, NaN or overflowed pivot
      rD[tid] = 1.0 / sqrt(dg);
    }
    if (bad && *info == 0) *info = j0 + 1;
    __syncthreads();
    for (int e = tid; e < kNB * kNB; e += kPiThreads) {
      const int r = e >> 7, c = e & (kNB - 1);
      if (r < nb && c <= r) S[r * st + c] = (c == r) ? S[r * st + r] * rD[r] : S[r * st + c] * rD[c];
    }
    __syncthreads();
  }
#if defined(PROBE_STOP) && PROBE_STOP == 1
  return;
#endif
  // ---- inverse, diagonal sub-blocks
  if (warp < nsub) {
    const int cI = 32 * warp;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      double a = (i == lane) ? 1.0 : 0.0;
      const double* li = S + (cI + i) * st + cI;
#pragma unroll
      for (int k = 0; k < i; ++k) a = fma(-li[k], x[k], a);  // x[k] == 0 for k < lane
      x[i] = (i >= lane) ? a * rD[cI + i] : 0.0;
    }
#pragma unroll
    for (int i = 1; i < 32; ++i)
      if (i > lane) S[(cI + lane) * st + cI + i] = x[i];  // X[cI + i][cI + lane]
  }
  __syncthreads();
#if defined(PROBE_STOP) && PROBE_STOP == 2
  return;
#endif
  // ---- inverse, off-diagonal sub-blocks by block distance: T_IJ = sum_{K = J}^{I - 1} L_IK X_KJ, X_IJ = -X_II T_IJ.
  // 64 threads per 32 x 32 block, each with a 4 x 4 register tile (rows tr + 8 u, columns tc + 8 v: lanes walk
  // consecutive rows / columns of the padded array, conflict-free): 8 shared-memory reads per 16 FMAs.  One output per
  // thread (3 reads per FMA) was shared-memory bound at ~23 us for the whole phase.
  for (int dlt = 1; dlt < nsub; ++dlt) {
    const int J = tid >> 6, tr = tid & 7, tc = (tid & 63) >> 3;
    const bool act = J < nsub - dlt;
    const int cI = 32 * (J + dlt), cJ = 32 * J;
    double acc[4][4];
    if (act) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
      double rdc[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) rdc[v] = rD[cJ + tc + 8 * v];
      // K = J: X_JJ is lower triangular -- X[kk][col] = 0 above the diagonal, 1 / L[col][col] on it
#pragma unroll 2
      for (int kk = cJ; kk < cJ + 32; ++kk) {
        double l[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) l[u] = S[(cI + tr + 8 * u) * st + kk];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int col = cJ + tc + 8 * v;
          const double raw = S[col * st + kk];
          x[v] = kk > col ? raw : (kk == col ? rdc[v] : 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = fma(l[u], x[v], acc[u][v]);
      }
#pragma unroll 2
      for (int kk = cJ + 32; kk < cI; ++kk) {
        double l[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) l[u] = S[(cI + tr + 8 * u) * st + kk];
#pragma unroll
        for (int v = 0; v < 4; ++v) x[v] = S[(cJ + tc + 8 * v) * st + kk];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = fma(l[u], x[v], acc[u][v]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) Tb[J * (32 * 33) + (tc + 8 * v) * 33 + tr + 8 * u] = acc[u][v];
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
      double rdr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rdr[u] = rD[cI + tr + 8 * u];
#pragma unroll 2
      for (int k = 0; k < 32; ++k) {
        double xi[4], t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = tr + 8 * u;
          const double raw = S[(cI + k) * st + cI + r];  // X[cI + r][cI + k] for k < r
          xi[u] = k < r ? raw : (k == r ? rdr[u] : 0.0);
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) t[v] = Tb[J * (32 * 33) + (tc + 8 * v) * 33 + k];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = fma(xi[u], t[v], acc[u][v]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) S[(cJ + tc + 8 * v) * st + cI + tr + 8 * u] = -acc[u][v];  // X[cI + r][cJ + c]
    }
    __syncthreads();
  }
#if defined(PROBE_STOP) && PROBE_STOP == 3
  return;
#endif
  for (int e = tid; e < kNB * kNB; e += kPiThreads) {
    const int i = e >> 7, c = e & (kNB - 1);
    const bool in = i < nb && c < nb;
    if (in && c <= i) A[(size_t)(j0 + i) * ld + j0 + c] = S[i * st + c];
    double v = 0.0, vt = 0.0;
    if (in) {
      if (c < i) v = S[c * st + i];        // Linv[i][c] = X[i][c]
      else if (c > i) vt = S[i * st + c];  // LinvT[i][c] = X[c][i]
      else v = vt = rD[i];
    }
    Linv[e] = v;
    LinvT[e] = vt;
  }
}

// after the panel GEMM Tn = -L21 (rows x nb, leading dimension kNB): A21 <- L21 and T21 <- L21^T (nb x rows, leading
// dimension ldt), the B operand of the trailing updates
__global__ void finish_panel_kernel(const double* __restrict__ Tn, int rows, int nb, double* __restrict__ A21, int ld,
                                    double* __restrict__ T21, int ldt) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int r = r0 + dy, c = c0 + threadIdx.x;
    double v = 0.0;
    if (r < rows && c < nb) {
      v = Tn[(size_t)r * kNB + c];
      A21[(size_t)r * ld + c] = -v;
    }
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int c = c0 + dy, r = r0 + threadIdx.x;
    if (c < nb && r < rows) T21[(size_t)c * ldt + r] = -tile[threadIdx.x][dy];
  }
}

// out[c] = sum_i Q[i][c] v[i]  (c < tp); one CTA of 1024 threads
__global__ void __launch_bounds__(1024) qt_vec_kernel(const double* __restrict__ Q, int n, int tp,
                                                      const double* __restrict__ v, double* __restrict__ out) {
  __shared__ double red[1024];
  const int c = threadIdx.x % tp, lane_rows = blockDim.x / tp, rr = threadIdx.x / tp;
  double s = 0.0;
  if (rr < lane_rows)
    for (int i = rr; i < n; i += lane_rows) s = fma(Q[(size_t)i * tp + c], v[i], s);
  red[threadIdx.x] = (rr < lane_rows) ? s : 0.0;
  __syncthreads();
  if (threadIdx.x < tp) {
    double tot = 0.0;
    for (int q = 0; q < lane_rows; ++q) tot += red[q * tp + threadIdx.x];
    out[threadIdx.x] = tot;
  }
}
// v[i] -= sum_c Q[i][c] q[c]; rows >= n of a padded vector are set to zero
__global__ void sub_q_kernel(const double* __restrict__ Q, int n, int tp, const double* __restrict__ q,
                             double* __restrict__ v, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  if (i >= n) { v[i] = 0.0; return; }
  double s = 0.0;
  for (int c = 0; c < tp; ++c) s = fma(Q[(size_t)i * tp + c], q[c], s);
  v[i] -= s;
}
// c_out[0 .. t) = 0 and c_out[t + i] = v[i] - sum_c Q[i][c] q[c] for i < n: the projected weights in the coefficient layout
__global__ void project_out_kernel(const double* __restrict__ Q, int n, int tp, int t, const double* __restrict__ q,
                                   const double* __restrict__ v, double* __restrict__ c_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < t) c_out[i] = 0.0;
  if (i >= n) return;
  double s = 0.0;
  for (int c = 0; c < tp; ++c) s = fma(Q[(size_t)i * tp + c], q[c], s);
  c_out[t + i] = v[i] - s;
}
__global__ void copy_pad_kernel(const double* __restrict__ src, int n, double* __restrict__ dst, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < n ? src[i] : 0.0;
}
__global__ void neg_kernel(const double* __restrict__ a, double* __restrict__ b, long long n) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) b[e] = -a[e];
}
int gram(b200rbf_ctx* h, const double* X, const double* Y, int n, int tp, double* partial, double* G) {
  const int nparts = (n + 63) / 64;
  const size_t smem = 2 * 64 * (size_t)tp * sizeof(double);
  B2_CUDA(cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_partial_kernel<<<nparts, 256, smem, h->stream>>>(X, Y, n, tp, partial);
  LAUNCHED(h);
  gram_reduce_kernel<<<(tp * tp + 255) / 256, 256, 0, h->stream>>>(partial, nparts, tp, G);
  LAUNCHED(h);
  return 0;
}

}  // namespace

// c_out (t + n) <- [0; RBF weights]; keeps Q, R in the workspace for the tail step.  info_dev: 0 ok, != 0 -> fall back to LU.
// cap_rows > n: the factor, Q / W, the diagonal-block inverses and the vectors are laid out for cap_rows rows so that
// fit_extend.cu can append points to this factorisation in place (no regrowth, no copies).
int chol_fit_weights(b200rbf_ctx* h, const double* Xu_dev, const double* f_dev, int n, int d, int kernel, double eta,
                     CholWork* wk, int* info_dev, int cap_rows, double* c_out) {
  FitState& F = h->fit;
  const int t = d + 1, tp = (t + 15) & ~15, n16 = (n + 31) & ~31;  // n16: padded to the 32-column sub-blocks
  const int rcap = cap_rows > n16 ? ((cap_rows + kNB - 1) / kNB) * kNB : n16, ld = rcap;
  const int nparts = (n + 63) / 64;
  size_t need = 0;
  auto take = [&](size_t cnt) { size_t off = need; need += (cnt + 15) & ~(size_t)15; return off; };
  const size_t oP = take((size_t)rcap * tp), oQn = take((size_t)n16 * tp), oW = take((size_t)rcap * tp);
  const size_t oAc = take((size_t)n16 * 3 * tp), oBc = take((size_t)3 * tp * ld), oT21 = take((size_t)2 * kOuter * ld);
  const size_t oG = take((size_t)tp * tp), oR1 = take((size_t)tp * tp);
  const size_t oR2 = take((size_t)tp * tp), oR = take((size_t)tp * tp), oS = take((size_t)tp * tp);
  const size_t oGp = take((size_t)nparts * tp * tp), oV0 = take(rcap), oV1 = take(rcap), oV2 = take(rcap), oq = take(tp);
  const size_t oLi = take((size_t)((rcap + kNB - 1) / kNB) * kNB * kNB), oLiT = take((size_t)kNB * kNB);
  const size_t oTn = take((size_t)n16 * kNB);
  const size_t oRi = take((size_t)tp * tp), oFq = take(tp), oBv = take(rcap), oU = take(rcap), oMr = take(rcap), oGb = take(rcap);
  const size_t oPt = take((size_t)(rcap / 256 + 2) * 2 * tp);
  B2_TRY(F.chol.reserve(need * sizeof(double)));
  B2_TRY(F.A.reserve(((size_t)rcap * ld + 1024) * sizeof(double)));
  double* base = F.chol.as<double>();
  CholWork& w = *wk;
  w.P = base + oP; w.Qneg = base + oQn; w.W = base + oW; w.Acat = base + oAc; w.Bcat = base + oBc; w.T21 = base + oT21;
  w.G = base + oG; w.R1 = base + oR1; w.R2 = base + oR2; w.R = base + oR; w.S = base + oS;
  w.Linv = base + oLi; w.LinvT = base + oLiT; w.Tn = base + oTn; w.gpart = base + oGp; w.v0 = base + oV0; w.v1 = base + oV1; w.v2 = base + oV2; w.q = base + oq;
  w.Rinv = base + oRi; w.fq = base + oFq; w.bvec = base + oBv; w.u = base + oU; w.mrow = base + oMr; w.gbuf = base + oGb; w.part = base + oPt;
  w.n = n; w.n16 = n16; w.t = t; w.tp = tp; w.ld = ld;
  double* A = F.A.as<double>();
  double* Q = w.P;  // P is orthonormalised in place

  // 1. Phi + eta I (pad columns zero)
  B2_TRY(assemble_plain(h, Xu_dev, n, d, kernel, eta, A, ld, n16));
  // 2. P and its thin QR by CholeskyQR2
  {
    const long long tot = (long long)n16 * tp;
    build_p_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(Xu_dev, n, d, t, w.P, tp, n16);
    LAUNCHED(h);
  }
  const size_t sm_small = (size_t)tp * tp * sizeof(double);
  B2_CUDA(cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
  B2_CUDA(cudaFuncSetAttribute(apply_rinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
  B2_CUDA(cudaFuncSetAttribute(build_acat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
  for (int pass = 0; pass < 2; ++pass) {
    double* Rp = pass == 0 ? w.R1 : w.R2;
    B2_TRY(gram(h, Q, Q, n, tp, w.gpart, w.G));
    chol_small_kernel<<<1, 256, sm_small, h->stream>>>(w.G, tp, t, Rp, info_dev);
    LAUNCHED(h);
    apply_rinv_kernel<<<(n + 127) / 128, 128, sm_small, h->stream>>>(Q, n, tp, t, Rp);
    LAUNCHED(h);
  }
  small_mm_kernel<<<(tp * tp + 255) / 256, 256, 0, h->stream>>>(w.R2, w.R1, w.R, tp);  // P = Q (R2 R1)
  LAUNCHED(h);
  // 3. W = (Phi + eta I) Q  as  W = 0 - A (-Q)
  {
    const long long tot = (long long)n16 * tp;
    // Qneg = -Q (pad rows are zero already), W = 0
    B2_CUDA(cudaMemsetAsync(w.W, 0, (size_t)tot * sizeof(double), h->stream));
    neg_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(Q, w.Qneg, tot);
    LAUNCHED(h);
    B2_TRY(gemm_sub_ex(h, w.W, tp, A, ld, w.Qneg, tp, n, tp, n16, 0));
  }
  // 4. S = Q^T W;  M = A - Acat Bcat on the lower triangle
  B2_TRY(gram(h, Q, w.W, n, tp, w.gpart, w.S));
  const double sigma = 1.0;
  build_acat_kernel<<<(n + 127) / 128, 128, sm_small, h->stream>>>(Q, w.W, w.S, n, tp, t, sigma, w.Acat);
  LAUNCHED(h);
  {
    dim3 blk(32, 8), grd((tp + 31) / 32, (n + 31) / 32);
    transpose_kernel<<<grd, blk, 0, h->stream>>>(w.W, n, tp, tp, w.Bcat, ld);                       // W^T
    LAUNCHED(h);
    transpose_kernel<<<grd, blk, 0, h->stream>>>(Q, n, tp, tp, w.Bcat + (size_t)tp * ld, ld);       // Q^T
    LAUNCHED(h);
    transpose_kernel<<<grd, blk, 0, h->stream>>>(Q, n, tp, tp, w.Bcat + (size_t)2 * tp * ld, ld);   // Q^T
    LAUNCHED(h);
  }
  B2_TRY(gemm_sub_ex(h, A, ld, w.Acat, 3 * tp, w.Bcat, ld, n, n, 3 * tp, 1));
  if (n16 > n) {
    const int cnt = (n16 - n) * n16;
    pad_identity_kernel<<<(cnt + 255) / 256, 256, 0, h->stream>>>(A, ld, n, n16);
    LAUNCHED(h);
  }
  // 5. blocked Cholesky of the n16 x n16 lower triangle, two levels.  Per 128-column step: potrf_inv_kernel factors
  // the diagonal block and inverts its factor; the panel below becomes ONE DMMA GEMM  Tn = 0 - A21 L11^-T  (no
  // sequential triangular solve over the long dimension); finish_panel_kernel stores L21 and L21^T; only the columns
  // that remain inside the current kOuter-wide outer panel are updated right away.  Everything to the right of the
  // outer panel is updated once per outer panel with K = kOuter: the DMMA GEMM runs at 32.8 TFLOP/s with K = 512
  // against 27.3 with K = 128 (C is read and written once per 512 instead of once per 128 columns).
  // The inverses of the diagonal blocks are kept for step 6.
  const size_t sm_pi = ((size_t)kNB * (kNB + 1) + 3 * 32 * 33) * sizeof(double);
  B2_CUDA(cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pi));
  // Look-ahead across outer panels.  The trailing update of outer panel J is split into C1 (the 512 columns of the NEXT
  // outer panel) and C2 (everything right of those).  As soon as C1 is done the next panel's chain -- 4 x (one-CTA
  // factorisation + inverse, panel GEMM, store, in-panel update: ~1 ms of launches that cannot fill the GPU) -- runs on
  // the high-priority stream WHILE C2, one big K = 512 GEMM, keeps the SMs busy on the main stream: the panel's CTAs
  // take the slots that retiring GEMM CTAs free.  L21^T is double-buffered (C2 of panel J reads buffer J & 1 while panel
  // J + 1 writes the other).  The first version ran everything in one stream: 157 x 75 us of potrf_inv_kernel alone
  // (10 % of the fit at N = 20 000) with 147 SMs idle.
  {
    cudaStream_t s_main = h->stream, s_la = h->la_stream;
    struct StreamGuard {  // every error return puts the handle's stream back
      b200rbf_ctx* h;
      cudaStream_t s;
      ~StreamGuard() { h->stream = s; }
    } stream_guard{h, s_main};
    B2_CUDA(cudaEventRecord(h->ev_la1, s_main));  // the assembled matrix
    B2_CUDA(cudaStreamWaitEvent(s_la, h->ev_la1, 0));
    int pj = 0;
    for (int J0 = 0; J0 < n16; J0 += kOuter, ++pj) {
      const int R0 = (J0 + kOuter < n16) ? J0 + kOuter : n16;  // first column right of the outer panel
      double* T21b = w.T21 + (size_t)(pj & 1) * kOuter * ld;
      h->stream = s_la;
      for (int j0 = J0; j0 < R0; j0 += kNB) {
        const int nb = (n16 - j0 < kNB) ? (n16 - j0) : kNB;
        const int r0 = j0 + nb, rows = n16 - r0;
        potrf_inv_kernel<<<1, kPiThreads, sm_pi, h->stream>>>(A, ld, j0, nb, w.Linv + (size_t)(j0 / kNB) * kNB * kNB,
                                                              w.LinvT, info_dev);
        LAUNCHED(h);
        if (rows <= 0) break;
        double* A21 = A + (size_t)r0 * ld + j0;
        double* T21 = T21b + (size_t)(j0 - J0) * ld + r0;  // T21[k][r - r0] = L[r][j0 + k]
        B2_CUDA(cudaMemsetAsync(w.Tn, 0, (size_t)rows * kNB * sizeof(double), h->stream));
        B2_TRY(gemm_sub_ex(h, w.Tn, kNB, A21, ld, w.LinvT, kNB, rows, nb, nb, 0));
        dim3 blk(32, 8), grd((nb + 31) / 32, (rows + 31) / 32);
        finish_panel_kernel<<<grd, blk, 0, h->stream>>>(w.Tn, rows, nb, A21, ld, T21, ld);
        LAUNCHED(h);
        if (r0 < R0)  // the rest of the outer panel: rows >= r0, columns [r0, R0)
          B2_TRY(gemm_sub_ex(h, A + (size_t)r0 * ld + r0, ld, A21, ld, T21, ld, rows, R0 - r0, nb, 1));
      }
      B2_CUDA(cudaEventRecord(h->ev_la2, s_la));  // panel pj factored, its L21^T complete
      h->stream = s_main;
      B2_CUDA(cudaStreamWaitEvent(s_main, h->ev_la2, 0));
      if (R0 < n16) {
        const int R1 = (R0 + kOuter < n16) ? R0 + kOuter : n16;
        // C1: the next outer panel's columns [R0, R1), rows >= R0
        B2_TRY(gemm_sub_ex(h, A + (size_t)R0 * ld + R0, ld, A + (size_t)R0 * ld + J0, ld, T21b + R0, ld, n16 - R0, R1 - R0,
                           R0 - J0, 1));
        B2_CUDA(cudaEventRecord(h->ev_la1, s_main));
        B2_CUDA(cudaStreamWaitEvent(s_la, h->ev_la1, 0));  // the next panel may start; C2 of this one overlaps it
        if (R1 < n16)  // C2: everything right of the next panel, K = R0 - J0
          B2_TRY(gemm_sub_ex(h, A + (size_t)R1 * ld + R1, ld, A + (size_t)R1 * ld + J0, ld, T21b + R1, ld, n16 - R1, n16 - R1,
                             R0 - J0, 1));
      }
    }
  }
  // 6. c = Pi L^-T L^-1 Pi f : two single-launch pipelined triangular solves (fit_extend.cu) with the stored inverses
  // of the diagonal blocks.  bvec = Pi f and u = L^-1 Pi f are kept: an extension appends to them.
  copy_pad_kernel<<<(n16 + 255) / 256, 256, 0, h->stream>>>(f_dev, n, w.bvec, n16);
  LAUNCHED(h);
  qt_vec_kernel<<<1, 1024, 0, h->stream>>>(Q, n, tp, w.bvec, w.fq);
  LAUNCHED(h);
  sub_q_kernel<<<(n16 + 255) / 256, 256, 0, h->stream>>>(Q, n, tp, w.fq, w.bvec, n16);
  LAUNCHED(h);
  B2_TRY(trsv_pipelined(h, A, ld, w.Linv, n, w.bvec, w.u, false));
  B2_TRY(trsv_pipelined(h, A, ld, w.Linv, n, w.u, w.v2, true));
  qt_vec_kernel<<<1, 1024, 0, h->stream>>>(Q, n, tp, w.v2, w.q);
  LAUNCHED(h);
  project_out_kernel<<<((n > t ? n : t) + 255) / 256, 256, 0, h->stream>>>(Q, n, tp, t, w.q, w.v2, c_out);
  LAUNCHED(h);
  B2_TRY(ext_make_rinv(h, w));  // R^-1: the tail (chol_fit_tail) and every later extension use it
  return 0;  // RBF weights in c_out[t .. t + n), c_out[0 .. t) = 0
}

// lambda and the final coefficient vector: g = Phi c (without eta) evaluated by the caller
// Tail coefficients in ONE single-CTA launch:  lambda = R^-1 Q0^T (f_0 - ((Phi + eta I) c)_0).  Because
// Q0^T (Phi + eta I)_{0,:} = W^T with W = [(Phi00 + eta I) Q0 ; rows Q0^T phi(X0, x_i) of the appended points] -- kept by
// the fit (step 3) and the extension -- the projected residual is  Q0^T f_0 - W^T c = fq - W^T c:  an n x t reduction
// instead of re-evaluating the n x n kernel matrix (the first version ran the fused score kernel over all centers for
// this).  Then lambda = R^-1 (fq - W^T c) with the stored R^-1 (t <= 160).
// c_full = [lambda; c] where c already sits at c_full + t;  tailc_out (optional): the resident model's tail.
__global__ void __launch_bounds__(1024) tail_fused_kernel(const double* __restrict__ fq, const double* __restrict__ W,
                                                          int n_rows, int tp, int t, const double* __restrict__ Rinv,
                                                          double* c_full, double* tailc_out) {
  __shared__ double red[1024];
  extern __shared__ double lam[];  // tp values + tp right-hand side
  double* v = lam + tp;
  const double* c = c_full + t;
  const int tid = threadIdx.x;
  {
    const int a = tid % tp, groups = 1024 / tp, gr = tid / tp;
    double s = 0.0;
    if (gr < groups)
      for (int i = gr; i < n_rows; i += groups) s = fma(W[(size_t)i * tp + a], c[i], s);
    red[tid] = (gr < groups) ? s : 0.0;
    __syncthreads();
    if (tid < tp) {
      double tot = 0.0;
      for (int q = 0; q < groups; ++q) tot += red[q * tp + tid];
      v[tid] = fq[tid] - tot;
    }
    __syncthreads();
  }
  // lambda = R^-1 v with the stored inverse (upper triangular, kept for the extension): one row per thread.  The first
  // version back-substituted column by column, two barriers and a dependent global load per unknown: 16 us at t = 31
  // of a 155 us extension.
  if (tid < t) {
    double s = 0.0;
    for (int k = tid; k < t; ++k) s = fma(Rinv[tid * tp + k], v[k], s);
    lam[tid] = s;
  }
  __syncthreads();
  for (int i = tid; i < t; i += 1024) {
    c_full[i] = lam[i];
    if (tailc_out != nullptr) tailc_out[i] = lam[i];
  }
}

// n_rows: all points carrying a weight (rows of W); c_full_dev + t holds their weights
int chol_fit_tail(b200rbf_ctx* h, const CholWork& w, double* c_full_dev, double* tailc_out, int n_rows) {
  tail_fused_kernel<<<1, 1024, 2 * (size_t)w.tp * sizeof(double), h->stream>>>(w.fq, w.W, n_rows, w.tp, w.t, w.Rinv, c_full_dev,
                                                                              tailc_out);
  LAUNCHED(h);
  return 0;
}

}  // namespace b200rbf
