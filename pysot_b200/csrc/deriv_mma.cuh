// K3x (many queries): gradient of the RBF interpolant on the FP64 tensor path.
//
//   grad_k(u) = sum_j s_j (u_k - c_jk) = u_k sum_j s_j  -  (S C)_k,        s_j = phi'(r_j) w_j / r_j     (rbf.py:209-213)
//
// so the d scaled accumulations per pair of deriv_tile_kernel become a second GEMM, S (m x n) times the centers C (n x d),
// chained behind the norm-expansion GEMM of the score kernel (score_mma.cuh) inside one warp, tile by tile:
//   GEMM 1  D = U C^T on mma.sync.m8n8k4.f64 (K = d): r^2 = |u|^2 + |c|^2 - 2 u.c (AUG: straight from the MMA), near pairs
//           recomputed directly, exactly as in score_mma_kernel;
//   scalar  s = phi'(r) w / r from r^2 -- no division, no square root for the thin-plate spline (deriv_tile.cuh);
//   GEMM 2  G += S C on the same instruction (K = the tile's centers, N = d padded to 8).
// The D fragment of GEMM 1 is the A fragment of GEMM 2 WITHOUT any shuffle: thread (gid, tig) of an m8n8 block owns
// columns 2 tig, 2 tig + 1; GEMM 2 takes its four contraction indices per k-step as {0, 2, 4, 6} + e, so slot tig is
// column 2 tig + e -- the thread's own value.  The eight centers of a block are assigned to the columns through the
// permutation pi = (0 1 2 3 5 4 7 6), which makes BOTH B-fragment access patterns (rows pi(gid), columns tig for GEMM 1;
// rows pi(2 tig + e), columns gid for GEMM 2) bank-conflict free with the row stride ds == 4 (mod 16) doubles.
// The gradient accumulators (16 candidates x d per warp) stay in registers for the whole sweep over the centers; the
// first version kept them in shared memory and touched them once per 4 centers.
// Work per pair: 4 K4 + 8 ND multiply-adds on the tensor path (ND = ceil(K4 / 2) n8 blocks of dimensions) + 8 .. 13 scalar
// FP64 operations, against 4 d + 7 .. 12 scalar operations before: 40 vs 52 at d = 10, 124 vs 212 at d = 50.
// Cubic and thin-plate-spline kernels only: for the linear kernel s = w / r is unbounded at coincident points and
// u_k s - s c_jk would cancel catastrophically there; it keeps deriv_tile_kernel.
#pragma once
#include "deriv_tile.cuh"

namespace b200rbf {

struct DerivMmaArgs {
  const double* x;      // m x d queries, original coordinates
  long long m;
  int d;
  const double* lb;     // d
  const double* range;  // d : ub - lb
  const double* pts;    // npad x ds centers (unit box; the augmented copy when AUG)
  const double* cnorm;  // npad squared norms
  const double* coef;   // npad weights (pad 0)
  const double* tailc;  // ntail
  int ds, npad, tail;
  double* out;          // m x d
  const double2* tps_table;
};

__device__ __forceinline__ int dm_pi(int c) { return c < 4 ? c : (c ^ 1); }  // (0 1 2 3 5 4 7 6)

template <int K4, int KERN, bool AUG>
__global__ void __launch_bounds__(kMmaThreads, (K4 >= 9 ? 2 : 3)) deriv_mma_kernel(DerivMmaArgs a) {
  constexpr int ND = (K4 + 1) / 2;  // n8 blocks of gradient dimensions
  const int ds = a.ds;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_pts = reinterpret_cast<double*>(smem_raw);                    // [stages][rows * ds] (+ 8 doubles of slack)
  double* s_wts = s_pts + (size_t)kMmaStages * kMmaTileRows * ds + 8;      // [stages][rows]
  double* s_nrm = s_wts + kMmaStages * kMmaTileRows;                       // [stages][rows]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_nrm + kMmaStages * kMmaTileRows);
  const double2* s_tps = reinterpret_cast<const double2*>(s_full + kMmaStages + (kMmaStages & 1));
  if (KERN == B200RBF_KERNEL_TPS) {
    double2* dst = const_cast<double2*>(s_tps);
    for (int i = threadIdx.x; i < kTpsLogEntries; i += kMmaThreads) dst[i] = a.tps_table[i];
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const long long base = ((long long)blockIdx.x * (kMmaThreads / 32) + warp) * (8 * kMmaMB);
  const int ntiles = (a.npad + kMmaTileRows - 1) / kMmaTileRows;
  if (tid == 0) {
    for (int s = 0; s < kMmaStages; ++s) mbar_init(&s_full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % kMmaStages;
    const int r0 = t * kMmaTileRows;
    const int rows = min(kMmaTileRows, a.npad - r0);
    mbar_expect_tx(&s_full[s], (uint32_t)rows * (uint32_t)(ds * 8 + 16));
    tma_bulk_g2s(s_pts + (size_t)s * kMmaTileRows * ds, a.pts + (size_t)r0 * ds, (uint32_t)rows * ds * 8u, &s_full[s]);
    tma_bulk_g2s(s_wts + s * kMmaTileRows, a.coef + r0, rows * 8u, &s_full[s]);
    tma_bulk_g2s(s_nrm + s * kMmaTileRows, a.cnorm + r0, rows * 8u, &s_full[s]);
  };
  if (tid == 0)
    for (int t = 0; t < kMmaStages && t < ntiles; ++t) issue(t);

  // A fragments of GEMM 1: unit-box coordinates of queries base + gid + 8 i, dimensions 4 k + tig (as score_mma_kernel)
  double afr[kMmaMB][K4];
  double nx[kMmaMB];
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    const long long row = base + gid + 8 * i;
    const double* src = a.x + (row < a.m ? row : a.m - 1) * (long long)a.d;
    const bool live = row < a.m;  // rows past the end sit at the unit-box origin (no repeated near-pair work)
    double n2 = 0.0;
#pragma unroll
    for (int k = 0; k < K4; ++k) {
      const int dim = 4 * k + tig;
      double v = 0.0;
      if (dim < a.d && live) v = __ddiv_rn(__dsub_rn(__ldg(src + dim), __ldg(a.lb + dim)), __ldg(a.range + dim));  // utils.py:33
      afr[i][k] = AUG ? -2.0 * v : v;
      n2 = fma(v, v, n2);
    }
    n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
    n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
    nx[i] = n2;
    if (AUG) {
#pragma unroll
      for (int k = 0; k < K4; ++k) {
        if (4 * k + tig == a.d) afr[i][k] = n2;
        if (4 * k + tig == a.d + 1) afr[i][k] = 1.0;
      }
    }
  }
  double G[kMmaMB][ND][2];
  double ssum[kMmaMB];
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    ssum[i] = 0.0;
#pragma unroll
    for (int nb = 0; nb < ND; ++nb) G[i][nb][0] = G[i][nb][1] = 0.0;
  }
  // direct recomputation only for tiny pairs, r^2 < 2^-T (|u|^2 + |c|^2): the expansion's absolute error E in r^2 enters
  // a cubic gradient term as 1.5 w E (u - c)_k / r <= 1.5 w E whatever r is (T = 40: duplicates and negative r^2 only),
  // a thin-plate one as w E (u - c)_k / r^2 <= w E / r (T = 16: 3e-13 w at the threshold); score_mma.cuh has the argument
  constexpr int kNearShiftHi = (KERN == B200RBF_KERNEL_TPS ? 16 : 40) << 20;
  const int swap = tig >> 1;              // pi swaps the two columns of the quads tig = 2, 3
  const int rowA = dm_pi(gid);            // center row (within an 8-block) behind GEMM 1's column gid

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kMmaStages;
    mbar_wait(&s_full[s], (uint32_t)((t / kMmaStages) & 1));
    const double* tp = s_pts + (size_t)s * kMmaTileRows * ds;
    const double* tw = s_wts + s * kMmaTileRows;
    const double* tn = s_nrm + s * kMmaTileRows;
    double acc[kMmaMB][kMmaJB][2];
#pragma unroll
    for (int i = 0; i < kMmaMB; ++i)
#pragma unroll
      for (int jb = 0; jb < kMmaJB; ++jb) acc[i][jb][0] = acc[i][jb][1] = 0.0;
    {
      const double* pb = tp + rowA * ds + tig;
#pragma unroll
      for (int k = 0; k < K4; ++k) {
        double bfr[kMmaJB];
#pragma unroll
        for (int jb = 0; jb < kMmaJB; ++jb) bfr[jb] = pb[8 * jb * ds + 4 * k];
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i)
#pragma unroll
          for (int jb = 0; jb < kMmaJB; ++jb) dmma884(acc[i][jb][0], acc[i][jb][1], afr[i][k], bfr[jb]);
      }
    }
    // r^2, near test (integer pipe, high words), rare direct recomputation -- as in score_mma_kernel; column e of this
    // thread is center 2 tig + (e ^ swap) of the block
    bool near_any = false;
#pragma unroll
    for (int jb = 0; jb < kMmaJB; ++jb) {
      const double2 n2 = *reinterpret_cast<const double2*>(tn + 8 * jb + 2 * tig);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double cn = (e ^ swap) ? n2.y : n2.x;
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i) {
          if (AUG) {
            near_any |= __double2hiint(acc[i][jb][e]) < max(__double2hiint(nx[i]), __double2hiint(cn)) - kNearShiftHi;
          } else {
            const double sq = nx[i] + cn;
            const double r2 = fma(-2.0, acc[i][jb][e], sq);
            near_any |= __double2hiint(r2) < __double2hiint(sq) - kNearShiftHi;
            acc[i][jb][e] = r2;
          }
        }
      }
    }
    // tiny pairs (rare: duplicates and negative r^2): direct differences, one thread per pair (mma_exact_r2)
    if (near_any) {
#pragma unroll
      for (int jb = 0; jb < kMmaJB; ++jb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jc = 8 * jb + 2 * tig + (e ^ swap);
#pragma unroll
          for (int i = 0; i < kMmaMB; ++i) {
            const int thr = AUG ? max(__double2hiint(nx[i]), __double2hiint(tn[jc])) : __double2hiint(nx[i] + tn[jc]);
            if (__double2hiint(acc[i][jb][e]) < thr - kNearShiftHi) {
              const long long row = base + gid + 8 * i;
              acc[i][jb][e] = mma_exact_r2(a.x + (row < a.m ? row : a.m - 1) * (long long)a.d, a.d, a.lb, a.range, tp + jc * ds);
            }
          }
        }
    }
    // s = phi'(r) w / r in place of r^2 (r clamped to eps, rbf.py:201; pad rows carry weight 0)
#pragma unroll
    for (int jb = 0; jb < kMmaJB; ++jb) {
      const double2 w2 = *reinterpret_cast<const double2*>(tw + 8 * jb + 2 * tig);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double w = (e ^ swap) ? w2.y : w2.x;
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i) {
          const double sv = dphi_c_over_r_from_r2<KERN>(acc[i][jb][e], w, s_tps);
          acc[i][jb][e] = sv;
          ssum[i] += sv;
        }
      }
    }
    // GEMM 2: G[cand][dim] += sum_j s[cand][j] C[j][dim]; k-step (jb, e) contracts centers pi(2 tig' + e), tig' = 0..3
#pragma unroll
    for (int jb = 0; jb < kMmaJB; ++jb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double* pc = tp + (8 * jb + 2 * tig + (e ^ swap)) * ds + gid;
#pragma unroll
        for (int nb = 0; nb < ND; ++nb) {
          const double b2 = pc[8 * nb];
#pragma unroll
          for (int i = 0; i < kMmaMB; ++i) dmma884(G[i][nb][0], G[i][nb][1], acc[i][jb][e], b2);
        }
      }
    __syncthreads();
    if (tid == 0 && t + kMmaStages < ntiles) issue(t + kMmaStages);
  }

  // grad_k = (dP_k . lambda + u_k sum_j s_j - G_k) / (ub_k - lb_k)   (rbf.py:207-215); this thread holds dimensions
  // 8 nb + 2 tig + e of queries gid + 8 i
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 1);
    ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 2);
    const long long row = base + gid + 8 * i;
    if (row >= a.m) continue;
    const double* src = a.x + row * (long long)a.d;
    double* o = a.out + row * (long long)a.d;
#pragma unroll
    for (int nb = 0; nb < ND; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * nb + 2 * tig + e;
        if (k < a.d) {
          const double rg = __ldg(a.range + k);
          const double uk = __ddiv_rn(__dsub_rn(__ldg(src + k), __ldg(a.lb + k)), rg);
          const double tl = (a.tail == B200RBF_TAIL_LINEAR) ? __ldg(a.tailc + 1 + k) : 0.0;
          o[k] = __ddiv_rn(tl + fma(uk, ssum[i], -G[i][nb][e]), rg);
        }
      }
  }
}

template <int K4, int KERN, bool AUG>
int launch_deriv_mma_inst(b200rbf_ctx* h, const DerivMmaArgs& a) {
  auto kern = deriv_mma_kernel<K4, KERN, AUG>;
  const size_t smem = ((size_t)kMmaStages * kMmaTileRows * a.ds + 8 + 2 * (size_t)kMmaStages * kMmaTileRows + kMmaStages + 1 +
                       (KERN == B200RBF_KERNEL_TPS ? 2 * kTpsLogEntries : 0)) * 8;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (a.m + kMmaRowsPerCta - 1) / kMmaRowsPerCta;
  kern<<<(unsigned)blocks, kMmaThreads, smem, h->stream>>>(a);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

template <int K4>
int launch_deriv_mma_k4(b200rbf_ctx* h, const DerivMmaArgs& a, int kernel, bool aug) {
  if (kernel == B200RBF_KERNEL_CUBIC)
    return aug ? launch_deriv_mma_inst<K4, B200RBF_KERNEL_CUBIC, true>(h, a) : launch_deriv_mma_inst<K4, B200RBF_KERNEL_CUBIC, false>(h, a);
  return aug ? launch_deriv_mma_inst<K4, B200RBF_KERNEL_TPS, true>(h, a) : launch_deriv_mma_inst<K4, B200RBF_KERNEL_TPS, false>(h, a);
}

int launch_deriv_mma_group0(b200rbf_ctx*, const DerivMmaArgs&, int k4, int kernel, bool aug);
int launch_deriv_mma_group1(b200rbf_ctx*, const DerivMmaArgs&, int k4, int kernel, bool aug);
int launch_deriv_mma_group2(b200rbf_ctx*, const DerivMmaArgs&, int k4, int kernel, bool aug);
int launch_deriv_mma_group3(b200rbf_ctx*, const DerivMmaArgs&, int k4, int kernel, bool aug);

}  // namespace b200rbf
