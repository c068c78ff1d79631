// K1: RBFInterpolant._fit, initial branch (rbf.py:110-130) and coefficient solve (rbf.py:164-168) on the device.
//
//   K1a assemble_kernel : A = [[0, P^T], [P, phi(cdist(Xu, Xu)) + eta I]] written once, row-major, with the
//                         right-hand side [0; rhs] as an extra column N (so L^-1 P rhs falls out of the
//                         factorisation and only the back substitution remains).
//   K1b blocked right-looking LU with partial pivoting (LAPACK dgetrf semantics, rbf.py:123):
//         panel_kernel  : cooperative launch, the panel's rows live in the CTAs' shared memory, one grid
//                         barrier per column (pivot search + row exchange through a small global mailbox)
//         laswp_kernel  : row interchanges on the columns outside the panel (rows are contiguous: coalesced)
//         trsm32_kernel : U12 = L11^-1 A12 in 32-row steps, one thread per column
//         gemm_kernel   : trailing update A22 -= L21 U12 on the FP64 tensor cores
//                         (mma.sync.m8n8k4.f64 -> SASS DMMA), cp.async double-buffered 128x128x16 tiles
//   K1c back substitution U x = y : strip GEMV + diagonal-block solve per 128-row block.
//
// Any backward-stable solver reproduces the reference's predictions to its own reorder noise;
// partial pivoting keeps the same algorithm class as the reference's dgetrf.
#include <cooperative_groups.h>

#include "common.cuh"

namespace b200rbf {
namespace {

__device__ __forceinline__ double phi_r2(int kern, double r2) {
  if (kern == B200RBF_KERNEL_CUBIC) return r2 * sqrt(r2);
  if (kern == B200RBF_KERNEL_TPS) {
    const double eps2 = 4.930380657631323783823303533017413935457e-32;
    const double c = fmax(r2, eps2);
    return 0.5 * c * log(c);
  }
  return sqrt(r2);
}

// ------------------------------------------------------------------------------------------------ K1a
constexpr int kAsmTile = 32;
constexpr int kAsmChunk = 32;

template <bool EXACT>
__global__ void __launch_bounds__(256) assemble_kernel(const double* __restrict__ Xu, int n, int d, int t,
                                                       int tail, int kern, double eta,
                                                       const double* __restrict__ rhs, double* __restrict__ A,
                                                       int ld, int rhs_col, int ncols) {
  // tile (bi, bj) of the N x (N+1) augmented matrix; 256 threads, 4 entries each (rows ty, ty+8, ty+16, ty+24)
  __shared__ double s_i[kAsmTile][kAsmChunk + 1];
  __shared__ double s_j[kAsmTile][kAsmChunk + 1];
  const int N = n + t;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int gi0 = blockIdx.y * kAsmTile, gj0 = blockIdx.x * kAsmTile;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < d; k0 += kAsmChunk) {
    const int kc = min(kAsmChunk, d - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < kAsmTile * kAsmChunk; e += 256) {
      const int r = e / kAsmChunk, k = e % kAsmChunk;
      const int pi = gi0 + r - t, pj = gj0 + r - t;
      s_i[r][k] = (k < kc && pi >= 0 && pi < n) ? Xu[(size_t)pi * d + k0 + k] : 0.0;
      s_j[r][k] = (k < kc && pj >= 0 && pj < n) ? Xu[(size_t)pj * d + k0 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty + 8 * q;
      double a = acc[q];
      for (int k = 0; k < kc; ++k) {
        const double df = s_i[r][k] - s_j[tx][k];
        a = EXACT ? __dadd_rn(a, __dmul_rn(df, df)) : fma(df, df, a);
      }
      acc[q] = a;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = gi0 + ty + 8 * q, j = gj0 + tx;
    if (i >= N || j >= ncols) continue;
    double v;
    if (j >= N) {
      // right-hand side [0; rhs] (rbf.py:165) at an even column; a gap column (N odd) is zero
      v = (j == rhs_col && i >= t) ? rhs[i - t] : 0.0;
    } else if (i < t && j < t) {
      v = 0.0;  // rbf.py:119 zeros(ntail, ntail)
    } else if (i < t) {
      v = (i == 0) ? 1.0 : Xu[(size_t)(j - t) * d + (i - 1)];  // P^T
    } else if (j < t) {
      v = (j == 0) ? 1.0 : Xu[(size_t)(i - t) * d + (j - 1)];  // P, tails/linear_tail.py:31
    } else {
      v = phi_r2(kern, acc[q]);
      if (i == j) v += eta;  // rbf.py:115
    }
    A[(size_t)i * ld + j] = v;
  }
  (void)tail;
}

// ------------------------------------------------------------------------------------------------ grid barrier
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1ULL);
    while (ld_acquire_u64(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ K1b panel
constexpr int kPanelThreads = 256;

struct PanelArgs {
  double* A;
  int ld;
  int N;        // matrix order
  int j0;       // first column / row of the panel
  int nb;       // panel width (<= NBMAX)
  int R;        // rows per CTA
  int* ipiv;    // N
  // mailbox, double buffered by step parity: [2][G] values, [2][G] rows, [2][G][nbmax] rows, [2][nbmax] diag row
  double* mb_val;
  int* mb_row;
  double* mb_rows;
  double* mb_diag;
  int nbmax;
  unsigned long long* barrier;
  unsigned long long bar_base;
  int* info;    // first exactly-zero pivot (1-based), 0 if none
};

__global__ void __launch_bounds__(kPanelThreads) panel_kernel(PanelArgs a) {
  extern __shared__ double sm[];
  const int nb = a.nb, st = a.nb | 1;  // odd stride: conflict-free column walks
  double* S = sm;                       // R x st
  double* s_prow = S + (size_t)a.R * st;  // nb
  __shared__ double s_rv[kPanelThreads / 32];
  __shared__ int s_ri[kPanelThreads / 32];
  __shared__ double s_best_v;
  __shared__ int s_best_i;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kPanelThreads / 32;
  const int G = gridDim.x, g = blockIdx.x;
  const int row0 = a.j0 + g * a.R;                       // first global row of this CTA's strip
  const int nrows = max(0, min(a.R, a.N - row0));        // rows actually owned
  // load strip (each row's nb doubles are contiguous)
  for (int e = tid; e < nrows * nb; e += kPanelThreads) {
    const int r = e / nb, c = e - r * nb;
    S[r * st + c] = a.A[(size_t)(row0 + r) * a.ld + a.j0 + c];
  }
  __syncthreads();

  for (int jj = 0; jj < nb; ++jj) {
    const int par = jj & 1;
    const int diag = a.j0 + jj;                 // global row that must receive the pivot
    const int first = max(0, diag - row0);      // first active local row
    // (a) local pivot candidate: max |S[r][jj]| over active rows, lowest row on ties (idamax)
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int r = first + tid; r < nrows; r += kPanelThreads) {
      double v = fabs(S[r * st + jj]);
      if (!(v >= 0.0)) v = 0.0;  // NaN never wins a pivot search; keeps every index in range
      if (v > bv || (v == bv && r < bi)) {
        bv = v;
        bi = r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_rv[warp] = bv;
      s_ri[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nwarps; ++w)
        if (s_rv[w] > bv || (s_rv[w] == bv && s_ri[w] < bi)) {
          bv = s_rv[w];
          bi = s_ri[w];
        }
      s_best_v = bv;
      s_best_i = bi;
      a.mb_val[par * G + g] = bv;
      a.mb_row[par * G + g] = (bv >= 0.0) ? row0 + bi : -1;
    }
    __syncthreads();
    // (b) publish the candidate row, and the diagonal row if it is ours
    if (s_best_v >= 0.0) {
      const int br = s_best_i;
      double* dst = a.mb_rows + ((size_t)par * G + g) * a.nbmax;
      for (int c = tid; c < nb; c += kPanelThreads) dst[c] = S[br * st + c];
    }
    const bool own_diag = (diag >= row0 && diag < row0 + nrows);
    if (own_diag) {
      double* dst = a.mb_diag + (size_t)par * a.nbmax;
      for (int c = tid; c < nb; c += kPanelThreads) dst[c] = S[(diag - row0) * st + c];
    }
    // (c) one grid-wide barrier per column
    grid_barrier(a.barrier, a.bar_base + (unsigned long long)(jj + 1) * G);
    // (d) every CTA picks the same winner
    bv = -1.0;
    bi = 0x7fffffff;
    for (int q = tid; q < G; q += kPanelThreads) {
      const double v = __ldcg(a.mb_val + par * G + q);
      const int r = __ldcg(a.mb_row + par * G + q);
      if (r >= 0 && (v > bv || (v == bv && r < bi))) {
        bv = v;
        bi = r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_rv[warp] = bv;
      s_ri[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nwarps; ++w)
        if (s_rv[w] > bv || (s_rv[w] == bv && s_ri[w] < bi)) {
          bv = s_rv[w];
          bi = s_ri[w];
        }
      s_best_v = bv;
      s_best_i = bi;
    }
    __syncthreads();
    const int p = s_best_i;               // global pivot row
    const int pg = (p - a.j0) / a.R;      // its owner
    const double* wrow = a.mb_rows + ((size_t)par * G + pg) * a.nbmax;
    for (int c = tid; c < nb; c += kPanelThreads) s_prow[c] = __ldcg(wrow + c);
    // (e) exchange: pivot row's slot receives the old diagonal row, diagonal slot receives the pivot row
    if (p != diag) {
      if (pg == g) {
        const double* drow = a.mb_diag + (size_t)par * a.nbmax;
        for (int c = tid; c < nb; c += kPanelThreads) S[(p - row0) * st + c] = __ldcg(drow + c);
      }
    }
    __syncthreads();
    if (p != diag && own_diag) {
      for (int c = tid; c < nb; c += kPanelThreads) S[(diag - row0) * st + c] = s_prow[c];
    }
    if (g == 0 && tid == 0) {
      a.ipiv[diag] = p;
      if (s_best_v == 0.0 && *a.info == 0) *a.info = diag + 1;
    }
    __syncthreads();
    // (f) scale the multipliers and apply the rank-1 update to the rest of the panel
    const double piv = s_prow[jj];
    if (piv != 0.0) {
      const int ufirst = max(0, diag + 1 - row0);
      for (int r = ufirst + warp; r < nrows; r += nwarps) {
        const double l = S[r * st + jj] / piv;
        for (int c = jj + 1 + lane; c < nb; c += 32) S[r * st + c] = fma(-l, s_prow[c], S[r * st + c]);
        __syncwarp();
        if (lane == 0) S[r * st + jj] = l;
      }
    }
    __syncthreads();
  }
  // write the factored strip back
  for (int e = tid; e < nrows * nb; e += kPanelThreads) {
    const int r = e / nb, c = e - r * nb;
    a.A[(size_t)(row0 + r) * a.ld + a.j0 + c] = S[r * st + c];
  }
}

// ------------------------------------------------------------------------------------------------ row interchanges
// columns [0, j0) and [j0 + nb, ncols): thread per column, swaps applied in pivot order
__global__ void laswp_kernel(double* __restrict__ A, int ld, int j0, int nb, int ncols, const int* __restrict__ ipiv) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= j0) c += nb;
  if (c >= ncols) return;
  for (int jj = 0; jj < nb; ++jj) {
    const int r = j0 + jj, p = ipiv[r];
    if (p != r) {
      const double x = A[(size_t)r * ld + c], y = A[(size_t)p * ld + c];
      A[(size_t)r * ld + c] = y;
      A[(size_t)p * ld + c] = x;
    }
  }
}

// ------------------------------------------------------------------------------------------------ 32-row TRSM step
// X[r0 .. r0+32, c0 .. c1) <- L^-1 X with L = unit lower 32 x 32 block at A[r0 .. r0+32, r0 .. r0+32)
__global__ void __launch_bounds__(128) trsm32_kernel(double* __restrict__ A, int ld, int r0, int rows, int c0, int c1) {
  __shared__ double sL[32][33];
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int i = e >> 5, k = e & 31;
    sL[i][k] = (i < rows && k < i) ? A[(size_t)(r0 + i) * ld + r0 + k] : 0.0;
  }
  __syncthreads();
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1) return;
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = (i < rows) ? A[(size_t)(r0 + i) * ld + c] : 0.0;
#pragma unroll
  for (int i = 1; i < 32; ++i) {
    double s = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-sL[i][k], x[k], s);
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < rows) A[(size_t)(r0 + i) * ld + c] = x[i];
}

// ------------------------------------------------------------------------------------------------ DMMA GEMM
// C[M x Nc] -= A[M x K] * B[K x Nc], all row-major.  K % 16 == 0; pointers / leading dims 16-byte aligned.
//
// CTA tile 128 x 64, 256 threads = 8 warps as 4 (M) x 2 (N), warp tile 32 x 32 = 4 x 4 m8n8k4 accumulators (64
// registers), so two CTAs fit one SM and one CTA's prologue / epilogue hides behind the other's main loop.
// Operands stream through a 3-stage cp.async ring with ONE barrier per k-step.  The C tile is loaded into the
// accumulators up front (its latency hides behind the first k-steps) and A fragments are negated on the way in,
// so acc = C - A B accumulates directly and the epilogue is store-only.
constexpr int kGemmBM = 128, kGemmBN = 64, kGemmThreads = 256;  // BK (16 or 32) is a template parameter
template <int BK> __host__ __device__ constexpr int gemm_lda() { return BK + 4; }  // 20 / 36 doubles (== 4 mod 16): conflict-free A fragments
constexpr int kGemmLdB = kGemmBN + 4;   // 68 doubles (68 mod 16 == 4): same for 4 k x 4 columns
template <int BK> __host__ __device__ constexpr int gemm_stage_doubles() { return kGemmBM * gemm_lda<BK>() + BK * kGemmLdB; }
template <int BK> constexpr size_t gemm_smem_bytes(int stages) { return (size_t)stages * gemm_stage_doubles<BK>() * sizeof(double); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int kGemmStages, int kGemmBK>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_sub_kernel(double* __restrict__ C, int ldc,
                                                                   const double* __restrict__ A, int lda,
                                                                   const double* __restrict__ B, int ldb, int M,
                                                                   int Nc, int K, int lower) {
  extern __shared__ __align__(16) double gsm[];
  constexpr int kGemmLdA = gemm_lda<kGemmBK>(), kGemmStageDoubles = gemm_stage_doubles<kGemmBK>();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps, warp tile 32 x 32
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  if (lower && n0 > m0 + kGemmBM - 1) return;  // symmetric update: tiles strictly above the diagonal are skipped
  const int gid = lane >> 2, tig = lane & 3;

  auto load_tiles = [&](int stage, int k0) {
    double* sA = gsm + stage * kGemmStageDoubles;
    double* sB = sA + kGemmBM * kGemmLdA;
    // A tile: 128 rows x 16 k = 128 x 8 chunks of 16 B; rows clamped (results of clamped rows are discarded)
#pragma unroll
    for (int it = 0; it < kGemmBK / 4; ++it) {  // 128 rows x BK/2 chunks of 16 B
      const int e = tid + it * kGemmThreads;
      const int r = e / (kGemmBK / 2), ch = e % (kGemmBK / 2);
      const int gr = min(m0 + r, M - 1);
      cp_async16(sA + r * kGemmLdA + ch * 2, A + (size_t)gr * lda + k0 + ch * 2);
    }
    // B tile: 16 k x 64 columns = 16 x 32 chunks; columns past Nc are clamped into the row (results discarded)
#pragma unroll
    for (int it = 0; it < kGemmBK / 8; ++it) {  // BK rows x 32 chunks
      const int e = tid + it * kGemmThreads;
      const int r = e >> 5, ch = e & 31;
      int gc = n0 + ch * 2;
      if (gc >= Nc) gc = (Nc - 1) & ~1;
      cp_async16(sB + r * kGemmLdB + ch * 2, B + (size_t)(k0 + r) * ldb + gc);
    }
  };

  const int nk = K / kGemmBK;
#pragma unroll
  for (int s = 0; s < kGemmStages - 1; ++s) {
    if (s < nk) load_tiles(s, s * kGemmBK);
    cp_async_commit();
  }

  // acc <- C (thread holds C[row = gid][col = 2*tig, 2*tig+1] of each 8x8 tile); out-of-range entries are never stored
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm * 32 + i * 8 + gid;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn * 32 + j * 8 + tig * 2;
      acc[i][j][0] = acc[i][j][1] = 0.0;
      if (r < M) {
        const double* cp = C + (size_t)r * ldc + c;
        if (c + 1 < Nc) {
          const double2 v = *reinterpret_cast<const double2*>(cp);
          acc[i][j][0] = v.x;
          acc[i][j][1] = v.y;
        } else if (c < Nc) {
          acc[i][j][0] = cp[0];
        }
      }
    }
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kGemmStages - 2>();  // this thread's copies of tile kt have landed
    __syncthreads();                   // ... everyone's have, and everyone is done with the stage refilled below
    {
      const int nxt = kt + kGemmStages - 1;
      if (nxt < nk) load_tiles(nxt % kGemmStages, nxt * kGemmBK);
      cp_async_commit();
    }
    const double* sA = gsm + (kt % kGemmStages) * kGemmStageDoubles;
    const double* sB = sA + kGemmBM * kGemmLdA;
    const double* pa = sA + (wm * 32 + gid) * kGemmLdA + tig;
    const double* pb = sB + tig * kGemmLdB + wn * 32 + gid;
#pragma unroll
    for (int ks = 0; ks < kGemmBK; ks += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = -pa[i * 8 * kGemmLdA + ks];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = pb[ks * kGemmLdB + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm * 32 + i * 8 + gid;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn * 32 + j * 8 + tig * 2;
      double* cp = C + (size_t)r * ldc + c;
      if (c + 1 < Nc) {
        *reinterpret_cast<double2*>(cp) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else if (c < Nc) {
        cp[0] = acc[i][j][0];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K1c back substitution
constexpr int kBsBlock = 128;

// y[r] -= sum_{c in [c0, c1)} U[r][c] * x[c]   for rows r in [r0, r0 + rows): one warp per row
__global__ void __launch_bounds__(256) strip_gemv_kernel(const double* __restrict__ A, int ld, int r0, int rows,
                                                         int c0, int c1, const double* __restrict__ x,
                                                         double* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const double* row = A + (size_t)(r0 + warp) * ld;
  double s = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) s = fma(row[c], x[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[r0 + warp] -= s;
}

// solve the upper-triangular diagonal block: x[r0 .. r0+rows) = U_bb^-1 y[r0 .. r0+rows); single CTA
__global__ void __launch_bounds__(kBsBlock) diag_upper_solve_kernel(const double* __restrict__ A, int ld, int r0,
                                                                     int rows, double* __restrict__ y,
                                                                     double* __restrict__ x, int* __restrict__ info) {
  extern __shared__ double su[];  // rows x (kBsBlock + 1)
  __shared__ double sx[kBsBlock];
  const int st = kBsBlock + 1, tid = threadIdx.x;
  for (int i = 0; i < rows; ++i)  // thread = column: coalesced row reads, no integer division
    if (tid < rows) su[i * st + tid] = A[(size_t)(r0 + i) * ld + r0 + tid];
  double v = (tid < rows) ? y[r0 + tid] : 0.0;
  __syncthreads();
  double rdiag = 0.0;  // reciprocals of the diagonal, all at once and off the critical path (as LAPACK's dtrsv does not,
  if (tid < rows) {    // but dgetf2 does for its pivots)
    const double dgl = su[tid * st + tid];
    if (dgl == 0.0) atomicCAS(info, 0, r0 + tid + 1);
    rdiag = 1.0 / dgl;
  }
  for (int k = rows - 1; k >= 0; --k) {
    if (tid == k) {
      v *= rdiag;
      sx[k] = v;
    }
    __syncthreads();
    if (tid < k) v = fma(-su[tid * st + k], sx[k], v);
  }
  if (tid < rows) x[r0 + tid] = v;
}

__global__ void extract_col_kernel(const double* __restrict__ A, int ld, int col, int N, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) y[i] = A[(size_t)i * ld + col];
}

// ------------------------------------------------------------------------------------------------ FP64 peak probes
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-9;
  const double m = 1.0 + 1e-12, c = 1e-15;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = seed;
  const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-15;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma_884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

// both at once: every warp interleaves 8 independent DFMA chains with 8 independent DMMA chains.  If the FP64
// tensor op had its own datapath the combined rate would exceed either peak; it does not.
__global__ void __launch_bounds__(256) dmix_peak_kernel(double* out, int iters, double seed) {
  double acc[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i][0] = acc[i][1] = seed;
    f[i] = seed + i + threadIdx.x * 1e-9;
  }
  const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-15, m = 1.0 + 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma_884(acc[i][0], acc[i][1], a, b);
      f[i] = fma(f[i], m, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + f[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace

// =================================================================================================== host side
template <int STAGES, int BK>
static int gemm_sub_inst(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M,
                         int Nc, int K, int lower) {
  static DeviceOnce attr;
  constexpr size_t smem = gemm_smem_bytes<BK>(STAGES);
  if (attr.pending(h->device)) {
    B2_CUDA(cudaFuncSetAttribute(gemm_sub_kernel<STAGES, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CUDA(cudaFuncSetAttribute(gemm_sub_kernel<STAGES, BK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared));
    attr.done(h->device);
  }
  dim3 grid((Nc + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM);
  gemm_sub_kernel<STAGES, BK><<<grid, kGemmThreads, smem, h->stream>>>(C, ldc, A, lda, B, ldb, M, Nc, K, lower);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

// C -= A B on the FP64 tensor cores; lower != 0: only tiles touching the lower triangle (C square, symmetric update)
int gemm_sub_ex(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int Nc,
                int K, int lower) {
  if (M <= 0 || Nc <= 0 || K <= 0) return 0;
  // gemm_stages: 2 = 2-stage BK 16 (58 KB: leaves room for a look-ahead panel CTA), 3 = 3-stage BK 16,
  // 4 = 2-stage BK 32 (109 KB, half the barriers) when K allows it
  if (h->gemm_stages == 4 && K % 32 == 0) return gemm_sub_inst<2, 32>(h, C, ldc, A, lda, B, ldb, M, Nc, K, lower);
  return h->gemm_stages == 2 ? gemm_sub_inst<2, 16>(h, C, ldc, A, lda, B, ldb, M, Nc, K, lower)
                             : gemm_sub_inst<3, 16>(h, C, ldc, A, lda, B, ldb, M, Nc, K, lower);
}
static int gemm_sub(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M,
                    int Nc, int K) {
  return gemm_sub_ex(h, C, ldc, A, lda, B, ldb, M, Nc, K, 0);
}

// Phi + eta I alone (n x n, no tail block), pad columns [n, ncols) zeroed: the operand of the null-space Cholesky fit
int assemble_plain(b200rbf_ctx* h, const double* Xu_dev, int n, int d, int kernel, double eta, double* A, int ld,
                   int ncols) {
  dim3 grid((ncols + kAsmTile - 1) / kAsmTile, (n + kAsmTile - 1) / kAsmTile);
  if (h->exact_dist)
    assemble_kernel<true><<<grid, 256, 0, h->stream>>>(Xu_dev, n, d, 0, 0, kernel, eta, nullptr, A, ld, -1, ncols);
  else
    assemble_kernel<false><<<grid, 256, 0, h->stream>>>(Xu_dev, n, d, 0, 0, kernel, eta, nullptr, A, ld, -1, ncols);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

// y[r] -= sum_{c in [c0, c1)} A[r][c] x[c] for rows [r0, r0 + rows)  (one warp per row; shared with fit_chol.cu)
int strip_gemv(b200rbf_ctx* h, const double* A, int ld, int r0, int rows, int c0, int c1, const double* x, double* y) {
  if (rows <= 0 || c1 <= c0) return 0;
  strip_gemv_kernel<<<(rows * 32 + 255) / 256, 256, 0, h->stream>>>(A, ld, r0, rows, c0, c1, x, y);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int lu_factor_device(b200rbf_ctx* h, double* A, int ld, int N, int rcol0, int ncols, int* ipiv, int* info_dev) {
  // blocked right-looking LU of the leading N x N block of the N x ncols row-major matrix.  Columns
  // [rcol0, ncols) (rcol0 even, >= N) ride along as right-hand sides; columns [N, rcol0) are an alignment gap.
  const int NB = h->lu_panel;
  FitState& F = h->fit;
  const int G_max = h->sm_count;
  // mailbox + barrier
  const size_t mb_doubles = 2 * (size_t)G_max + 2 * (size_t)G_max * NB + 2 * (size_t)NB;
  B2_TRY(F.work.reserve(mb_doubles * sizeof(double) + 2 * (size_t)G_max * sizeof(int) + 64));
  B2_TRY(F.flags.reserve(64));
  double* mb_val = F.work.as<double>();
  double* mb_rows = mb_val + 2 * (size_t)G_max;
  double* mb_diag = mb_rows + 2 * (size_t)G_max * NB;
  int* mb_row = reinterpret_cast<int*>(mb_diag + 2 * (size_t)NB);
  unsigned long long* barrier = F.flags.as<unsigned long long>();
  B2_CUDA(cudaMemsetAsync(barrier, 0, 64, h->stream));
  unsigned long long bar_base = 0;

  int max_smem = 0;
  B2_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  static DeviceOnce attr;
  if (attr.pending(h->device)) {
    B2_CUDA(cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024));
    attr.done(h->device);
  }

  // factor the panel starting at column j0 on `stream`
  auto launch_panel = [&](int j0, cudaStream_t stream) -> int {
    const int nb = (N - j0 < NB) ? (N - j0) : NB;
    const int prow = N - j0;
    const int st = nb | 1;
    // rows per CTA: at least 32, at most what fits in shared memory
    const int rmax = (int)(((size_t)max_smem - 2048 - (size_t)nb * 8) / ((size_t)st * 8));
    int G = (prow + 31) / 32;
    if (G > G_max) G = G_max;
    int R = (prow + G - 1) / G;
    B2_CHECK(R <= rmax, B200RBF_EINVAL,
             "LU panel does not fit in shared memory: %d rows x %d columns per CTA (max %d rows); lower "
             "B200RBF_OPT_LU_PANEL",
             R, nb, rmax);
    G = (prow + R - 1) / R;
    PanelArgs pa;
    pa.A = A; pa.ld = ld; pa.N = N; pa.j0 = j0; pa.nb = nb; pa.R = R; pa.ipiv = ipiv;
    pa.mb_val = mb_val; pa.mb_row = mb_row; pa.mb_rows = mb_rows; pa.mb_diag = mb_diag; pa.nbmax = NB;
    pa.barrier = barrier; pa.bar_base = bar_base; pa.info = info_dev;
    const size_t smem = ((size_t)R * st + nb) * sizeof(double);
    if (stream == h->la_stream) {
      // Look-ahead: a cooperative launch would wait for the whole GPU to drain, i.e. for the trailing update it is
      // supposed to overlap.  A plain launch on the high-priority stream lets the panel's CTAs (at most one per SM)
      // take SM slots as GEMM CTAs retire; the GEMM never waits on the panel, so every panel CTA becomes resident
      // and the grid barrier completes.
      panel_kernel<<<G, kPanelThreads, smem, stream>>>(pa);
      B2_CUDA(cudaGetLastError());
    } else {
      void* kargs[] = {&pa};
      B2_CUDA(cudaLaunchCooperativeKernel((void*)panel_kernel, dim3(G), dim3(kPanelThreads), kargs, smem, stream));
    }
    h->launches++;
    bar_base += (unsigned long long)nb * G;
    return 0;
  };
  // apply the factored panel (j0, nb) to columns [ca, cb): U12 = L11^-1 A12 in 32-row steps, then A22 -= L21 U12
  auto update_cols = [&](int j0, int nb, int ca, int cb) -> int {
    if (ca >= cb) return 0;
    for (int s0 = 0; s0 < nb; s0 += 32) {
      const int rows = (nb - s0 < 32) ? (nb - s0) : 32;
      trsm32_kernel<<<(cb - ca + 127) / 128, 128, 0, h->stream>>>(A, ld, j0 + s0, rows, ca, cb);
      h->launches++;
      B2_CUDA(cudaGetLastError());
      const int below = nb - (s0 + rows);
      if (below > 0) {
        // rows [j0+s0+rows, j0+nb) of A12 -= L11[those rows, s0 .. s0+rows) * X[s0 .. s0+rows)   (rows == 32 here)
        B2_TRY(gemm_sub(h, A + (size_t)(j0 + s0 + rows) * ld + ca, ld, A + (size_t)(j0 + s0 + rows) * ld + j0 + s0, ld,
                        A + (size_t)(j0 + s0) * ld + ca, ld, below, cb - ca, rows));
      }
    }
    const int c0 = j0 + nb;
    if (c0 < N) {  // trailing rows
      B2_TRY(gemm_sub(h, A + (size_t)c0 * ld + ca, ld, A + (size_t)c0 * ld + j0, ld, A + (size_t)j0 * ld + ca, ld,
                      N - c0, cb - ca, nb));
    }
    return 0;
  };

  const bool lookahead = h->lu_lookahead && h->la_stream != nullptr;
  struct StagesGuard {  // restores the GEMM variant on every return path, the early error returns included
    b200rbf_ctx* h;
    int saved;
    ~StagesGuard() { h->gemm_stages = saved; }
  } stages_guard{h, h->gemm_stages};
  if (lookahead && h->gemm_stages_auto) h->gemm_stages = 2;
  B2_TRY(launch_panel(0, h->stream));
  for (int j0 = 0; j0 < N; j0 += NB) {
    const int nb = (N - j0 < NB) ? (N - j0) : NB;
    // row interchanges of this panel on every column outside it (the panel itself was swapped in shared memory)
    if (ncols - nb > 0) {
      laswp_kernel<<<(ncols - nb + 255) / 256, 256, 0, h->stream>>>(A, ld, j0, nb, ncols, ipiv);
      h->launches++;
      B2_CUDA(cudaGetLastError());
    }
    const int c0 = j0 + nb;
    if (c0 >= N) {  // final panel: only the right-hand-side columns remain (they start at the even column rcol0)
      B2_TRY(update_cols(j0, nb, rcol0, ncols));
      break;
    }
    // non-final panels have nb == NB (even), so c0 and c0 + nb1 below are 16-byte aligned columns
    const int nb1 = (N - c0 < NB) ? (N - c0) : NB;
    const int rest = (c0 + nb1 < N) ? c0 + nb1 : rcol0;
    B2_TRY(update_cols(j0, nb, c0, c0 + nb1));  // the next panel's columns first
    if (lookahead) {
      // panel k+1 touches only A[c0:, c0:c0+nb1), ipiv[c0:c0+nb1) and the mailbox: disjoint from what the rest of
      // update k reads (L21, U12 right of c0+nb1) and writes (columns >= c0+nb1)
      B2_CUDA(cudaEventRecord(h->ev_la1, h->stream));
      B2_CUDA(cudaStreamWaitEvent(h->la_stream, h->ev_la1, 0));
      B2_TRY(launch_panel(c0, h->la_stream));
      B2_CUDA(cudaEventRecord(h->ev_la2, h->la_stream));
      B2_TRY(update_cols(j0, nb, rest, ncols));
      B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_la2, 0));
    } else {
      B2_TRY(update_cols(j0, nb, rest, ncols));
      B2_TRY(launch_panel(c0, h->stream));
    }
  }
  return 0;
}

int back_substitute(b200rbf_ctx* h, const double* A, int ld, int N, double* y, double* x, int* info_dev) {
  static DeviceOnce attr;
  const size_t smem = (size_t)kBsBlock * (kBsBlock + 1) * sizeof(double);
  if (attr.pending(h->device)) {
    B2_CUDA(cudaFuncSetAttribute(diag_upper_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.done(h->device);
  }
  const int nblk = (N + kBsBlock - 1) / kBsBlock;
  for (int b = nblk - 1; b >= 0; --b) {
    const int r0 = b * kBsBlock;
    const int rows = (N - r0 < kBsBlock) ? (N - r0) : kBsBlock;
    const int c0 = r0 + rows;
    if (c0 < N) {
      strip_gemv_kernel<<<(rows * 32 + 255) / 256, 256, 0, h->stream>>>(A, ld, r0, rows, c0, N, x, y);
      h->launches++;
      B2_CUDA(cudaGetLastError());
    }
    diag_upper_solve_kernel<<<1, kBsBlock, smem, h->stream>>>(A, ld, r0, rows, y, x, info_dev);
    h->launches++;
    B2_CUDA(cudaGetLastError());
  }
  return 0;
}

// Xu_dev: n x d, rhs_dev: n, c_dev: n + t (all device).  info: 0 ok, >0 singular pivot index
int run_fit(b200rbf_ctx* h, const double* Xu_dev, const double* rhs_dev, int n, int d, int kernel, int tail,
            double eta, double* c_dev, int* info) {
  const int t = (tail == B200RBF_TAIL_LINEAR) ? d + 1 : 1;
  const int N = n + t;
  const int rhs_col = (N + 1) & ~1;
  const int ncols = rhs_col + 1;
  const int ld = (ncols + 7) & ~7;
  FitState& F = h->fit;
  B2_TRY(F.A.reserve(((size_t)N * ld + 1024) * sizeof(double)));
  B2_TRY(F.piv.reserve((size_t)N * sizeof(int) + 64));
  B2_TRY(F.rhs.reserve((size_t)N * sizeof(double) + 64));
  double* A = F.A.as<double>();
  int* ipiv = F.piv.as<int>();
  int* info_dev = ipiv + N;
  B2_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), h->stream));
  {
    dim3 grid((ncols + kAsmTile - 1) / kAsmTile, (N + kAsmTile - 1) / kAsmTile);
    if (h->exact_dist)
      assemble_kernel<true><<<grid, 256, 0, h->stream>>>(Xu_dev, n, d, t, tail, kernel, eta, rhs_dev, A, ld, rhs_col, ncols);
    else
      assemble_kernel<false><<<grid, 256, 0, h->stream>>>(Xu_dev, n, d, t, tail, kernel, eta, rhs_dev, A, ld, rhs_col, ncols);
    h->launches++;
    B2_CUDA(cudaGetLastError());
  }
  B2_TRY(lu_factor_device(h, A, ld, N, rhs_col, ncols, ipiv, info_dev));
  double* y = F.rhs.as<double>();
  extract_col_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(A, ld, rhs_col, N, y);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  B2_TRY(back_substitute(h, A, ld, N, y, c_dev, info_dev));
  B2_CUDA(cudaMemcpyAsync(info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

// test hook: LU-factor a caller-supplied N x N row-major matrix with nrhs extra columns (device, in place)
int run_lu_debug(b200rbf_ctx* h, double* A_dev, int ld, int N, int rcol0, int ncols, int* ipiv_dev, int* info_dev) {
  return lu_factor_device(h, A_dev, ld, N, rcol0, ncols, ipiv_dev, info_dev);
}
int run_gemm_debug(b200rbf_ctx* h, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M,
                   int Nc, int K) {
  return gemm_sub(h, C, ldc, A, lda, B, ldb, M, Nc, K);
}

int run_fp64_peaks(b200rbf_ctx* h, double ms, double* out) {
  DevBuf sink;
  B2_TRY(sink.reserve(64));
  const int blocks = h->sm_count * 8;
  cudaEvent_t e0, e1;
  B2_CUDA(cudaEventCreate(&e0));
  B2_CUDA(cudaEventCreate(&e1));
  for (int which = 0; which < 3; ++which) {
    int iters = 2000;
    float t = 0.f;
    for (int rep = 0; rep < 3; ++rep) {  // calibrate, then run ~ms
      B2_CUDA(cudaEventRecord(e0, h->stream));
      if (which == 0)
        dfma_peak_kernel<<<blocks, 256, 0, h->stream>>>(sink.as<double>(), iters, 1.0);
      else if (which == 1)
        dmma_peak_kernel<<<blocks, 256, 0, h->stream>>>(sink.as<double>(), iters, 1.0);
      else
        dmix_peak_kernel<<<blocks, 256, 0, h->stream>>>(sink.as<double>(), iters, 1.0);
      h->launches++;
      B2_CUDA(cudaEventRecord(e1, h->stream));
      B2_CUDA(cudaEventSynchronize(e1));
      B2_CUDA(cudaEventElapsedTime(&t, e0, e1));
      if (rep < 2) {
        double scale = ms / (t > 1e-3 ? t : 1e-3);
        if (rep == 0 && scale > 50) scale = 50;
        iters = (int)(iters * scale);
        if (iters < 100) iters = 100;
      }
    }
    const double threads = (double)blocks * 256;
    double flops;
    if (which == 0)
      flops = threads * 8.0 * 2.0 * iters;             // 8 FMAs per iteration per thread
    else if (which == 1)
      flops = (threads / 32.0) * 8.0 * 512.0 * iters;  // 8 m8n8k4 MMAs (8*8*4*2 flops) per iteration per warp
    else
      flops = threads * 8.0 * 2.0 * iters + (threads / 32.0) * 8.0 * 512.0 * iters;
    out[which] = flops / (t * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  sink.release();
  return 0;
}

}  // namespace b200rbf
