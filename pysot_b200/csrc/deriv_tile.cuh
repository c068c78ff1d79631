// K3 (many queries): gradient of the RBF interpolant with the structure of the score kernel -- one thread owns one
// query point (coordinates in registers), center rows + weights stream through a ring of shared-memory tiles filled
// by TMA bulk copies, four centers are processed together.  Per center pair the thread makes two passes over the
// coordinates: (A) r^2 = sum (u_k - c_jk)^2, then the scalar s_j = phi'(r_j) c_j / r_j (r clamped to eps, rbf.py:201),
// (B) g_k += (u_k - c_jk) s_j  (rbf.py:209-213).  The d gradient accumulators live in shared memory ([k][thread]:
// conflict-free) and are touched once per 4 centers, so the kernel stays bound by the FP64 pipe: 4 d + ~25 fp64
// operations per pair.  deriv_kernels.cu keeps the slab-parallel kernel for small query counts, where a thread per
// query cannot fill the GPU.
#pragma once
#include "score_mma.cuh"

namespace b200rbf {

constexpr int kDtThreads = 128;
constexpr int kDtTileRows = 32;
constexpr int kDtStages = 2;

struct DerivTileArgs {
  const double* x;      // m x d queries, original coordinates
  long long m;
  int d;
  const double* lb;     // d
  const double* range;  // d : ub - lb
  const double* pts;    // npad x ds centers (unit box, zero padded to ds >= 4 K4)
  const double* coef;   // npad weights (pad 0)
  const double* tailc;  // ntail
  int ds, npad, tail;
  double* out;          // m x d
  const double2* tps_table;  // half-log table (score_mma.cuh), thin-plate spline only
};

// s = phi'(r) c / r from r^2 (rbf.py:212 with r clamped to eps, rbf.py:201), without a division and -- for the
// thin-plate spline -- without a square root:
//   cubic  (cubic_kernel.py:37)   3 r^2 c / r = 3 c sqrt(r2)                 Newton-corrected rsqrt seed (score_mma.cuh)
//   TPS    (tps_kernel.py:41-42)  r (1 + 2 log r) c / r = c (1 + 2 * 0.5 log r2)    table-driven half-log
//   linear (linear_kernel.py:40)  c / r = c rsqrt(max(r2, eps^2))            two Newton steps on the hardware seed
// Below the clamp the cubic value differs from the reference's by < 1e-15 |c| and multiplies (u - c) < eps: invisible.
template <int KERN>
__device__ __forceinline__ double dphi_c_over_r_from_r2(double r2, double c, const double2* __restrict__ tps_tab) {
  if (KERN == B200RBF_KERNEL_CUBIC) {
    return (3.0 * c) * mma_phi_from_r2<B200RBF_KERNEL_LINEAR>(r2, nullptr);  // sqrt(r2), exact 0 at r2 = 0
  } else if (KERN == B200RBF_KERNEL_TPS) {
    const double cl = __hiloint2double(max(__double2hiint(r2), 0x39700000), __double2loint(r2));  // >= eps^2 = 2^-104
    return fma(2.0 * c, tps_half_log(cl, tps_tab), c);
  } else {
    const double x = fmax(r2, 4.930380657631323783823303533017413935457e-32);  // eps^2
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const double g = x * y;
      y = fma(y, fma(-g, 0.5 * y, 0.5), y);  // y <- y + y (1 - x y^2) / 2
    }
    return c * y;
  }
}

template <int K4, int KERN>
__global__ void __launch_bounds__(kDtThreads, 2) deriv_tile_kernel(DerivTileArgs a) {
  constexpr int DP = 4 * K4;
  const int ds = a.ds;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_pts = reinterpret_cast<double*>(smem_raw);                  // [stages][rows * ds]
  double* s_wts = s_pts + (size_t)kDtStages * kDtTileRows * ds;          // [stages][rows]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_wts + kDtStages * kDtTileRows);
  double* s_g = reinterpret_cast<double*>(s_full + kDtStages);          // [DP][threads]
  const double2* s_tps = reinterpret_cast<const double2*>(s_g + (size_t)DP * kDtThreads);  // TPS: half-log table
  if (KERN == B200RBF_KERNEL_TPS)
    for (int i = threadIdx.x; i < kTpsLogEntries; i += kDtThreads) const_cast<double2*>(s_tps)[i] = a.tps_table[i];

  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * kDtThreads + tid;
  const bool live = row < a.m;
  const int ntiles = (a.npad + kDtTileRows - 1) / kDtTileRows;
  if (tid == 0) {
    for (int s = 0; s < kDtStages; ++s) mbar_init(&s_full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % kDtStages;
    const int r0 = t * kDtTileRows;
    const int rows = min(kDtTileRows, a.npad - r0);
    mbar_expect_tx(&s_full[s], (uint32_t)rows * (uint32_t)(ds * 8 + 8));
    tma_bulk_g2s(s_pts + (size_t)s * kDtTileRows * ds, a.pts + (size_t)r0 * ds, (uint32_t)rows * ds * 8u, &s_full[s]);
    tma_bulk_g2s(s_wts + s * kDtTileRows, a.coef + r0, rows * 8u, &s_full[s]);
  };
  if (tid == 0)
    for (int t = 0; t < kDtStages && t < ntiles; ++t) issue(t);

  double u[DP];
  {
    const double* src = a.x + (live ? row : 0) * (long long)a.d;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      double v = 0.0;
      if (k < a.d) v = __ddiv_rn(__dsub_rn(__ldg(src + k), __ldg(a.lb + k)), __ldg(a.range + k));  // utils.py:33
      u[k] = v;
      s_g[k * kDtThreads + tid] = 0.0;
    }
  }
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kDtStages;
    const int rows = min(kDtTileRows, a.npad - t * kDtTileRows);  // multiple of 16
    mbar_wait(&s_full[s], (uint32_t)((t / kDtStages) & 1));
    const double* tp = s_pts + (size_t)s * kDtTileRows * ds;
    const double* tw = s_wts + s * kDtTileRows;
    for (int j0 = 0; j0 < rows; j0 += 4) {
      const double* cp = tp + j0 * ds;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < DP; k += 2) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const double2 c2 = *reinterpret_cast<const double2*>(cp + jj * ds + k);
          const double t0 = u[k] - c2.x, t1 = u[k + 1] - c2.y;
          acc[jj] = fma(t0, t0, acc[jj]);
          acc[jj] = fma(t1, t1, acc[jj]);
        }
      }
      double fac[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        fac[jj] = dphi_c_over_r_from_r2<KERN>(acc[jj], tw[j0 + jj], s_tps);  // pad rows carry weight 0
      }
#pragma unroll
      for (int k = 0; k < DP; k += 2) {
        double g0 = s_g[k * kDtThreads + tid], g1 = s_g[(k + 1) * kDtThreads + tid];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const double2 c2 = *reinterpret_cast<const double2*>(cp + jj * ds + k);
          g0 = fma(u[k] - c2.x, fac[jj], g0);
          g1 = fma(u[k + 1] - c2.y, fac[jj], g1);
        }
        s_g[k * kDtThreads + tid] = g0;
        s_g[(k + 1) * kDtThreads + tid] = g1;
      }
    }
    __syncthreads();
    if (tid == 0 && t + kDtStages < ntiles) issue(t + kDtStages);
  }
  if (live) {
    double* o = a.out + row * (long long)a.d;
    for (int k = 0; k < a.d; ++k) {
      const double g = (a.tail == B200RBF_TAIL_LINEAR) ? __ldg(a.tailc + 1 + k) : 0.0;  // dP . lambda (rbf.py:207-208)
      o[k] = __ddiv_rn(g + s_g[k * kDtThreads + tid], __ldg(a.range + k));             // rbf.py:215
    }
  }
}

template <int K4, int KERN>
int launch_deriv_tile_inst(b200rbf_ctx* h, const DerivTileArgs& a) {
  auto kern = deriv_tile_kernel<K4, KERN>;
  const size_t smem = ((size_t)kDtStages * kDtTileRows * a.ds + kDtStages * kDtTileRows + kDtStages +
                       (size_t)4 * K4 * kDtThreads + (KERN == B200RBF_KERNEL_TPS ? 2 * kTpsLogEntries : 0)) * 8;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((a.m + kDtThreads - 1) / kDtThreads), kDtThreads, smem, h->stream>>>(a);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

template <int K4>
int launch_deriv_tile_k4(b200rbf_ctx* h, const DerivTileArgs& a, int kernel) {
  switch (kernel) {
    case B200RBF_KERNEL_CUBIC: return launch_deriv_tile_inst<K4, B200RBF_KERNEL_CUBIC>(h, a);
    case B200RBF_KERNEL_TPS: return launch_deriv_tile_inst<K4, B200RBF_KERNEL_TPS>(h, a);
    default: return launch_deriv_tile_inst<K4, B200RBF_KERNEL_LINEAR>(h, a);
  }
}

int launch_deriv_tile_group0(b200rbf_ctx*, const DerivTileArgs&, int k4, int kernel);
int launch_deriv_tile_group1(b200rbf_ctx*, const DerivTileArgs&, int k4, int kernel);
int launch_deriv_tile_group2(b200rbf_ctx*, const DerivTileArgs&, int k4, int kernel);
int launch_deriv_tile_group3(b200rbf_ctx*, const DerivTileArgs&, int k4, int kernel);

}  // namespace b200rbf
