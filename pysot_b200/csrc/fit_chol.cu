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
//   1. Phi + eta I -> A (n16 x ld, n16 = n rounded to 16)                                    assemble_plain
//   2. P = [1, Xu] (n x tp, tp = t rounded to 16), Q, R by CholeskyQR2                       gram / chol_small / apply_rinv
//   3. W = (Phi + eta I) Q                                                                   DMMA GEMM (K = n16)
//   4. S = Q^T W,  M = A - [Q | W | -Q (S + sigma I)] [W^T ; Q^T ; Q^T]  (lower triangle)    DMMA GEMM (K = 3 tp)
//   5. blocked right-looking Cholesky, 128-column panels: diagonal block in one CTA's shared memory, 32-column
//      triangular solves + DMMA GEMM for the panel, DMMA GEMM (lower tiles only) for the trailing update
//   6. c = Pi L^-T L^-1 Pi f   (strip GEMV + 128 x 128 triangular solves)
//   7. lambda from step 7 of the header formula; (Phi + eta I) c is evaluated by the fused score kernel (api.cu)
#include "common.cuh"

namespace b200rbf {

int gemm_sub_ex(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int Nc,
                int K, int lower);
int assemble_plain(b200rbf_ctx* h, const double* Xu_dev, int n, int d, int kernel, double eta, double* A, int ld,
                   int ncols);
int strip_gemv(b200rbf_ctx* h, const double* A, int ld, int r0, int rows, int c0, int c1, const double* x, double* y);

namespace {

constexpr int kNB = 128;

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

// Cholesky of the nb x nb diagonal block at (j0, j0) in one CTA; writes L back (lower) and L^T into Lt (nb x kNB).
// Right-looking, thread = (row i, column residue mod 8): three barriers per column, conflict-free shared-memory
// walks (row stride nb + 1), no integer division in the update loop.
constexpr int kPotrfThreads = 1024, kPotrfParts = kPotrfThreads / kNB;
__global__ void __launch_bounds__(kPotrfThreads) potrf_kernel(double* __restrict__ A, int ld, int j0, int nb,
                                                    double* __restrict__ Lt, int* __restrict__ info) {
  extern __shared__ double s[];  // nb x (nb + 1)
  const int st = nb + 1, tid = threadIdx.x;
  const int col = tid & (kNB - 1), half = tid >> 7;  // kNB == 128; `half` = column residue class 0 .. kPotrfParts-1
  for (int i = half; i < nb; i += kPotrfParts)
    if (col < nb) s[i * st + col] = col <= i ? A[(size_t)(j0 + i) * ld + j0 + col] : 0.0;
  __syncthreads();
  const int i = col;  // this thread's row in the update phase
  for (int k = 0; k < nb; ++k) {
    if (tid == k) {
      const double dgl = s[k * st + k];
      if (!(dgl > 0.0) && *info == 0) *info = j0 + k + 1;
      s[k * st + k] = sqrt(dgl);
    }
    __syncthreads();
    if (half == 0 && i > k && i < nb) s[i * st + k] /= s[k * st + k];
    __syncthreads();
    if (i > k && i < nb) {
      const double lik = s[i * st + k];
      for (int j = k + 1 + half; j <= i; j += kPotrfParts) s[i * st + j] = fma(-lik, s[j * st + k], s[i * st + j]);
    }
    // the next column's pivot s[k+1][k+1] is final once row k+1's own thread(s) are done: the barrier at the top of
    // the next iteration (after the sqrt by thread k+1) orders it for everyone else
    __syncthreads();
  }
  for (int r = half; r < nb; r += kPotrfParts) {
    if (col < nb) {
      if (col <= r) A[(size_t)(j0 + r) * ld + j0 + col] = s[r * st + col];
      Lt[(size_t)r * kNB + col] = col >= r ? s[col * st + r] : 0.0;  // Lt[r][c] = L[c][r]
    }
  }
}

// X[r0 .. r0+nrows, c0 .. c0+w) <- X L^-T with L the lower w x w diagonal block at (c0, c0); one thread per row
__global__ void __launch_bounds__(128) trsm_right32_kernel(double* __restrict__ A, int ld, int r0, int nrows, int c0,
                                                           int w) {
  __shared__ double sL[32][33];
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int i = e >> 5, k = e & 31;
    sL[i][k] = (i < w && k <= i) ? A[(size_t)(c0 + i) * ld + c0 + k] : (i == k ? 1.0 : 0.0);
  }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double* row = A + (size_t)(r0 + r) * ld + c0;
  double x[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) x[c] = c < w ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    double v = x[c];
#pragma unroll
    for (int k = 0; k < c; ++k) v = fma(-x[k], sL[c][k], v);
    x[c] = v / sL[c][c];
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < w) row[c] = x[c];
}

// Inverses of ALL diagonal blocks of the lower factor in one launch (CTA b inverts the block at (b kNB, b kNB)):
// thread j builds column j of L^-1 by forward substitution out of shared memory.  With them the two triangular
// solves need no sequential in-block substitution: a block step is a 128 x 128 matrix-vector product.
__global__ void __launch_bounds__(kNB) block_inverse_kernel(const double* __restrict__ A, int ld, int n16,
                                                            double* __restrict__ Linv) {
  // one kNB x (kNB + 1) array: L in the lower triangle (diagonal included), X = L^-1 strictly below its diagonal is
  // kept TRANSPOSED in the upper triangle (X[i][j] at s[j][i], i > j), its diagonal in sD
  extern __shared__ double s[];
  __shared__ double sD[kNB];
  const int st = kNB + 1, tid = threadIdx.x;
  const int r0 = blockIdx.x * kNB, rows = (n16 - r0 < kNB) ? (n16 - r0) : kNB;
  for (int i = 0; i < rows; ++i)
    if (tid <= i) s[i * st + tid] = A[(size_t)(r0 + i) * ld + r0 + tid];
  __syncthreads();
  const int j = tid;
  if (j < rows) {
    const double xjj = 1.0 / s[j * st + j];
    sD[j] = xjj;
    double* xr = s + j * st;  // xr[k] = X[k][j] for k > j
    for (int i = j + 1; i < rows; ++i) {
      const double* li = s + i * st;
      double a0 = -li[j] * xjj, a1 = 0.0;
      int k = j + 1;
      for (; k + 1 < i; k += 2) {
        a0 = fma(-li[k], xr[k], a0);
        a1 = fma(-li[k + 1], xr[k + 1], a1);
      }
      if (k < i) a0 = fma(-li[k], xr[k], a0);
      xr[i] = (a0 + a1) / li[i];
    }
  }
  __syncthreads();
  double* out = Linv + (size_t)blockIdx.x * kNB * kNB;
  for (int i = 0; i < kNB; ++i) {
    double v = 0.0;
    if (i < rows && tid < rows) v = tid < i ? s[tid * st + i] : (tid == i ? sD[i] : 0.0);
    out[(size_t)i * kNB + tid] = v;
  }
}

// x[0 .. rows) = B y (transpose == 0) or B^T y (transpose != 0) for one kNB x kNB lower-triangular block B = L_bb^-1
__global__ void __launch_bounds__(kNB) block_matvec_kernel(const double* __restrict__ B, int rows,
                                                           const double* __restrict__ y, double* __restrict__ x,
                                                           int transpose) {
  extern __shared__ double s[];  // kNB x (kNB + 1)
  __shared__ double sy[kNB];
  const int st = kNB + 1, tid = threadIdx.x;
  sy[tid] = tid < rows ? y[tid] : 0.0;
  if (!transpose) {  // row dot products: stage the block so the row walks are conflict-free
    for (int i = 0; i < rows; ++i) s[i * st + tid] = B[(size_t)i * kNB + tid];
    __syncthreads();
    if (tid < rows) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int k = 0;
      for (; k + 3 <= tid; k += 4) {
        a0 = fma(s[tid * st + k], sy[k], a0);
        a1 = fma(s[tid * st + k + 1], sy[k + 1], a1);
        a2 = fma(s[tid * st + k + 2], sy[k + 2], a2);
        a3 = fma(s[tid * st + k + 3], sy[k + 3], a3);
      }
      for (; k <= tid; ++k) a0 = fma(s[tid * st + k], sy[k], a0);
      x[tid] = (a0 + a1) + (a2 + a3);
    }
  } else {  // column dot products: the global reads are already coalesced
    __syncthreads();
    if (tid < rows) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int i = tid;
      for (; i + 3 < rows; i += 4) {
        a0 = fma(B[(size_t)i * kNB + tid], sy[i], a0);
        a1 = fma(B[(size_t)(i + 1) * kNB + tid], sy[i + 1], a1);
        a2 = fma(B[(size_t)(i + 2) * kNB + tid], sy[i + 2], a2);
        a3 = fma(B[(size_t)(i + 3) * kNB + tid], sy[i + 3], a3);
      }
      for (; i < rows; ++i) a0 = fma(B[(size_t)i * kNB + tid], sy[i], a0);
      x[tid] = (a0 + a1) + (a2 + a3);
    }
  }
}

// y[c] -= sum_{r in [r0, r0+rows)} A[r][c] x[r]  for c < ccount  (thread per column: coalesced over c)
__global__ void col_update_kernel(const double* __restrict__ A, int ld, int r0, int rows, int ccount,
                                  const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double sx[kNB];
  if (threadIdx.x < rows) sx[threadIdx.x] = x[r0 + threadIdx.x];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ccount) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s = fma(A[(size_t)(r0 + r) * ld + c], sx[r], s);
  y[c] -= s;
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
__global__ void copy_pad_kernel(const double* __restrict__ src, int n, double* __restrict__ dst, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < n ? src[i] : 0.0;
}
// lambda = R^-1 q (R upper t x t, leading dimension tp), then c_full = [lambda; c]
__global__ void tail_finish_kernel(const double* __restrict__ R, int tp, int t, const double* __restrict__ q,
                                   const double* __restrict__ c, int n, double* __restrict__ c_full) {
  extern __shared__ double lam[];
  if (threadIdx.x == 0) {
    for (int i = t - 1; i >= 0; --i) {
      double v = q[i];
      for (int k = i + 1; k < t; ++k) v = fma(-R[i * tp + k], lam[k], v);
      lam[i] = v / R[i * tp + i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < t; i += blockDim.x) c_full[i] = lam[i];
  for (int i = threadIdx.x; i < n; i += blockDim.x) c_full[t + i] = c[i];
}
__global__ void neg_kernel(const double* __restrict__ a, double* __restrict__ b, long long n) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) b[e] = -a[e];
}
// r[i] = f[i] - (g[i] + eta c[i])
__global__ void residual_kernel(const double* __restrict__ f, const double* __restrict__ g,
                                const double* __restrict__ c, double eta, int n, double* __restrict__ r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) r[i] = f[i] - fma(eta, c[i], g[i]);
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

// c_rbf (n) <- RBF weights; keeps Q, R in the workspace for the tail step.  info_dev: 0 ok, != 0 -> fall back to LU.
int chol_fit_weights(b200rbf_ctx* h, const double* Xu_dev, const double* f_dev, int n, int d, int kernel, double eta,
                     CholWork* wk, int* info_dev) {
  FitState& F = h->fit;
  const int t = d + 1, tp = (t + 15) & ~15, n16 = (n + 15) & ~15, ld = n16;
  const int nparts = (n + 63) / 64;
  size_t need = 0;
  auto take = [&](size_t cnt) { size_t off = need; need += (cnt + 15) & ~(size_t)15; return off; };
  const size_t oP = take((size_t)n16 * tp), oQn = take((size_t)n16 * tp), oW = take((size_t)n16 * tp);
  const size_t oAc = take((size_t)n16 * 3 * tp), oBc = take((size_t)3 * tp * ld), oT21 = take((size_t)kNB * ld);
  const size_t oLt = take((size_t)kNB * kNB), oG = take((size_t)tp * tp), oR1 = take((size_t)tp * tp);
  const size_t oR2 = take((size_t)tp * tp), oR = take((size_t)tp * tp), oS = take((size_t)tp * tp);
  const size_t oGp = take((size_t)nparts * tp * tp), oV0 = take(n16), oV1 = take(n16), oV2 = take(n16), oq = take(tp);
  const size_t oLi = take((size_t)((n16 + kNB - 1) / kNB) * kNB * kNB);
  B2_TRY(F.chol.reserve(need * sizeof(double)));
  B2_TRY(F.A.reserve(((size_t)n16 * ld + 1024) * sizeof(double)));
  double* base = F.chol.as<double>();
  CholWork& w = *wk;
  w.P = base + oP; w.Qneg = base + oQn; w.W = base + oW; w.Acat = base + oAc; w.Bcat = base + oBc; w.T21 = base + oT21;
  w.Lt = base + oLt; w.G = base + oG; w.R1 = base + oR1; w.R2 = base + oR2; w.R = base + oR; w.S = base + oS;
  w.Linv = base + oLi; w.gpart = base + oGp; w.v0 = base + oV0; w.v1 = base + oV1; w.v2 = base + oV2; w.q = base + oq;
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
  // 5. blocked Cholesky of the n16 x n16 lower triangle
  const size_t sm_potrf = (size_t)kNB * (kNB + 1) * sizeof(double);
  B2_CUDA(cudaFuncSetAttribute(potrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_potrf));
  for (int j0 = 0; j0 < n16; j0 += kNB) {
    const int nb = (n16 - j0 < kNB) ? (n16 - j0) : kNB;
    potrf_kernel<<<1, kPotrfThreads, (size_t)nb * (nb + 1) * sizeof(double), h->stream>>>(A, ld, j0, nb, w.Lt, info_dev);
    LAUNCHED(h);
    const int r0 = j0 + nb, rows = n16 - r0;
    if (rows <= 0) break;
    for (int s0 = 0; s0 < nb; s0 += 32) {
      const int wdt = (nb - s0 < 32) ? (nb - s0) : 32;
      trsm_right32_kernel<<<(rows + 127) / 128, 128, 0, h->stream>>>(A, ld, r0, rows, j0 + s0, wdt);
      LAUNCHED(h);
      const int rest = nb - s0 - wdt;
      if (rest > 0)  // A21[:, s0+w:] -= X[:, s0:s0+w] * (L11[s0+w:, s0:s0+w])^T  with the transposed block taken from Lt
        B2_TRY(gemm_sub_ex(h, A + (size_t)r0 * ld + j0 + s0 + wdt, ld, A + (size_t)r0 * ld + j0 + s0, ld,
                           w.Lt + (size_t)s0 * kNB + s0 + wdt, kNB, rows, rest, wdt, 0));
    }
    {
      dim3 blk(32, 8), grd((nb + 31) / 32, (rows + 31) / 32);
      transpose_kernel<<<grd, blk, 0, h->stream>>>(A + (size_t)r0 * ld + j0, rows, nb, ld, w.T21, ld);  // L21^T
      LAUNCHED(h);
    }
    B2_TRY(gemm_sub_ex(h, A + (size_t)r0 * ld + r0, ld, A + (size_t)r0 * ld + j0, ld, w.T21, ld, rows, rows, nb, 1));
  }
  // 6. c = Pi L^-T L^-1 Pi f
  copy_pad_kernel<<<(n16 + 255) / 256, 256, 0, h->stream>>>(f_dev, n, w.v0, n16);
  LAUNCHED(h);
  qt_vec_kernel<<<1, 1024, 0, h->stream>>>(Q, n, tp, w.v0, w.q);
  LAUNCHED(h);
  sub_q_kernel<<<(n16 + 255) / 256, 256, 0, h->stream>>>(Q, n, tp, w.q, w.v0, n16);
  LAUNCHED(h);
  const size_t sm_blk = (size_t)kNB * (kNB + 1) * sizeof(double);
  B2_CUDA(cudaFuncSetAttribute(block_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_blk));
  B2_CUDA(cudaFuncSetAttribute(block_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_blk));
  const int nblk = (n16 + kNB - 1) / kNB;
  block_inverse_kernel<<<nblk, kNB, sm_blk, h->stream>>>(A, ld, n16, w.Linv);
  LAUNCHED(h);
  for (int b = 0; b < nblk; ++b) {  // forward: L y = b   (y in v1), right-looking: every row below updates in parallel
    const int r0 = b * kNB, rows = (n16 - r0 < kNB) ? (n16 - r0) : kNB;
    block_matvec_kernel<<<1, kNB, sm_blk, h->stream>>>(w.Linv + (size_t)b * kNB * kNB, rows, w.v0 + r0, w.v1 + r0, 0);
    LAUNCHED(h);
    B2_TRY(strip_gemv(h, A, ld, r0 + rows, n16 - r0 - rows, r0, r0 + rows, w.v1, w.v0));
  }
  for (int b = nblk - 1; b >= 0; --b) {  // backward: L^T x = y   (x in v2, y consumed in v1)
    const int r0 = b * kNB, rows = (n16 - r0 < kNB) ? (n16 - r0) : kNB;
    block_matvec_kernel<<<1, kNB, sm_blk, h->stream>>>(w.Linv + (size_t)b * kNB * kNB, rows, w.v1 + r0, w.v2 + r0, 1);
    LAUNCHED(h);
    if (r0 > 0) {
      col_update_kernel<<<(r0 + 255) / 256, 256, 0, h->stream>>>(A, ld, r0, rows, r0, w.v2, w.v1);
      LAUNCHED(h);
    }
  }
  qt_vec_kernel<<<1, 1024, 0, h->stream>>>(Q, n, tp, w.v2, w.q);
  LAUNCHED(h);
  sub_q_kernel<<<(n16 + 255) / 256, 256, 0, h->stream>>>(Q, n, tp, w.q, w.v2, n16);
  LAUNCHED(h);
  return 0;  // RBF weights in w.v2[0 .. n)
}

// lambda and the final coefficient vector: g = Phi c (without eta) evaluated by the caller
int chol_fit_tail(b200rbf_ctx* h, const CholWork& w, const double* f_dev, const double* g_dev, double eta,
                  double* c_full_dev) {
  residual_kernel<<<(w.n + 255) / 256, 256, 0, h->stream>>>(f_dev, g_dev, w.v2, eta, w.n, w.v0);
  LAUNCHED(h);
  qt_vec_kernel<<<1, 1024, 0, h->stream>>>(w.P, w.n, w.tp, w.v0, w.q);
  LAUNCHED(h);
  tail_finish_kernel<<<1, 256, (size_t)w.tp * sizeof(double), h->stream>>>(w.R, w.tp, w.t, w.q, w.v2, w.n, c_full_dev);
  LAUNCHED(h);
  return 0;
}

}  // namespace b200rbf
