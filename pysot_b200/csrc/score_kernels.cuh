// Fused candidate x center kernel (K2) and its distance-only sibling (K2d), sm_100a.
//
// Replaces, for one tile of candidates, the reference's
//   cdist(to_unit_box(cand), _X) -> kernel.eval -> np.dot(., c) + np.dot(tail.eval(.), lambda)   rbf.py:180-185
//   cdist(cand, vstack(X, Xpend)) -> np.amin(axis=1)                                             candidate_srbf.py:35-36
//   x.min(), x.max() of both results                                                             utils.py:60-61
// without ever storing an m x n matrix.
//
// Work decomposition: one thread owns one candidate; its (padded) coordinates live in registers
// (DPAD is a template parameter so the coordinate loop is fully unrolled).  Center rows stream through a
// kStages-deep ring of shared-memory tiles filled by TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx); every lane reads the same center element, so the LDS.128 are pure broadcasts.  kJT centers
// are processed together to give the FP64 pipe independent dependency chains.  The kernel is bound by the
// FP64 FMA pipe: per pair DPAD subtractions + DPAD FMAs + sqrt/phi/accumulate.
#pragma once
#include "common.cuh"

namespace b200rbf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <int KERN>
__device__ __forceinline__ double rbf_phi_from_r2(double r2) {
  // kernels/cubic_kernel.py:25 r^3 ; tps_kernel.py:28-29 r clamped to eps then r^2 log r ; linear_kernel.py:28 r
  if (KERN == B200RBF_KERNEL_CUBIC) {
    return r2 * sqrt(r2);
  } else if (KERN == B200RBF_KERNEL_TPS) {
    const double eps2 = 4.930380657631323783823303533017413935457e-32;  // eps^2
    double c = fmax(r2, eps2);
    return 0.5 * c * log(c);  // r^2 log r = 0.5 r^2 log r^2
  } else {
    return sqrt(r2);
  }
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct TileSmem {
  // layout inside dynamic shared memory: [kStages][kTileRows*DPAD] points, [kStages][kTileRows] weights, barriers
  double* pts;
  double* wts;
  uint64_t* full;
};

template <int DPAD, int KERN, bool EXACT, bool WITH_PHI>
__global__ void __launch_bounds__(kScoreThreads, kScoreMinBlocks) score_kernel(ScoreArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_pts = reinterpret_cast<double*>(smem_raw);
  double* s_wts = s_pts + (size_t)kStages * kTileRows * DPAD;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_wts + kStages * kTileRows);
  double* s_red = reinterpret_cast<double*>(s_full + kStages);  // 4 x (kScoreThreads/32)

  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * kScoreThreads + tid;
  const bool live = row < a.m;
  const int ntiles = (a.npad + kTileRows - 1) / kTileRows;
  constexpr uint32_t kRowBytes = DPAD * sizeof(double);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int s = t % kStages;
    const int r0 = t * kTileRows;
    const int rows = min(kTileRows, a.npad - r0);
    uint32_t bytes = rows * kRowBytes + (WITH_PHI ? rows * 8u : 0u);
    mbar_expect_tx(&s_full[s], bytes);
    tma_bulk_g2s(s_pts + (size_t)s * kTileRows * DPAD, a.pts + (size_t)r0 * DPAD, rows * kRowBytes, &s_full[s]);
    if (WITH_PHI) tma_bulk_g2s(s_wts + s * kTileRows, a.coef + r0, rows * 8u, &s_full[s]);
  };
  if (tid == 0) {
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);
  }

  // candidate coordinates -> registers (unit-box map of utils.py:33 when lb is given)
  double u[DPAD];
  {
    const double* src = a.cand + (live ? row : 0) * (long long)a.d;
#pragma unroll
    for (int k = 0; k < DPAD; ++k) {
      double v = 0.0;
      if (k < a.d) {
        v = __ldg(src + k);
        if (a.lb != nullptr) v = __ddiv_rn(__dsub_rn(v, __ldg(a.lb + k)), __ldg(a.range + k));
      }
      u[k] = v;
    }
  }

  double sum = 0.0;
  double minr2 = __longlong_as_double(0x7ff0000000000000LL);

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    const int rows = min(kTileRows, a.npad - t * kTileRows);
    mbar_wait(&s_full[s], (uint32_t)((t / kStages) & 1));
    const double* tp = s_pts + (size_t)s * kTileRows * DPAD;
    const double* tw = s_wts + s * kTileRows;
    for (int j0 = 0; j0 < rows; j0 += kJT) {
      double acc[kJT];
#pragma unroll
      for (int jj = 0; jj < kJT; ++jj) acc[jj] = 0.0;
      const double* cp = tp + j0 * DPAD;
#pragma unroll
      for (int k = 0; k < DPAD; k += 2) {
#pragma unroll
        for (int jj = 0; jj < kJT; ++jj) {
          const double2 c2 = *reinterpret_cast<const double2*>(cp + jj * DPAD + k);
          const double t0 = u[k] - c2.x;
          const double t1 = u[k + 1] - c2.y;
          if (EXACT) {
            acc[jj] = __dadd_rn(acc[jj], __dmul_rn(t0, t0));
            acc[jj] = __dadd_rn(acc[jj], __dmul_rn(t1, t1));
          } else {
            acc[jj] = fma(t0, t0, acc[jj]);
            acc[jj] = fma(t1, t1, acc[jj]);
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < kJT; ++jj) {
        minr2 = fmin(minr2, acc[jj]);
        if (WITH_PHI) sum = fma(rbf_phi_from_r2<KERN>(acc[jj]), tw[j0 + jj], sum);
      }
    }
    __syncthreads();  // everyone is done with stage s: safe to refill it
    if (tid == 0 && t + kStages < ntiles) issue(t + kStages);
  }

  double fval = 0.0, dval = 0.0;
  if (WITH_PHI) {
    // tail: lambda_0 + sum_k lambda_k u_k  (tails/linear_tail.py:31, constant_tail.py:30; rbf.py:184)
    double tsum = __ldg(a.tailc);
    if (a.tail == B200RBF_TAIL_LINEAR) {
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < a.d) tsum = fma(__ldg(a.tailc + 1 + k), u[k], tsum);
    }
    fval = sum + tsum;
    if (live) a.fvals[row] = fval;
  }
  if (a.min_mode != 0) {
    dval = a.dscale * sqrt(minr2);
    if (live) {
      if (a.min_mode == 2) dval = fmin(dval, a.dmerit[row]);
      a.dmerit[row] = dval;
    }
  }

  // per-CTA min / max partials (unit_rescale statistics)
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double fmn = live ? fval : inf, fmx = live ? fval : -inf, dmn = live ? dval : inf, dmx = live ? dval : -inf;
  fmn = warp_min(fmn);
  fmx = warp_max(fmx);
  dmn = warp_min(dmn);
  dmx = warp_max(dmx);
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kScoreThreads / 32;
  if (lane == 0) {
    s_red[warp] = fmn;
    s_red[NW + warp] = fmx;
    s_red[2 * NW + warp] = dmn;
    s_red[3 * NW + warp] = dmx;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NW; ++w) {
      fmn = fmin(fmn, s_red[w]);
      fmx = fmax(fmx, s_red[NW + w]);
      dmn = fmin(dmn, s_red[2 * NW + w]);
      dmx = fmax(dmx, s_red[3 * NW + w]);
    }
    const int slot = a.part_off + blockIdx.x;
    if (WITH_PHI) {
      a.parts[slot] = fmn;
      a.parts[a.nparts + slot] = fmx;
    }
    if (a.min_mode != 0) {
      a.parts[2 * a.nparts + slot] = dmn;
      a.parts[3 * a.nparts + slot] = dmx;
    }
  }
}

template <int DPAD>
constexpr size_t score_smem_bytes() {
  return (size_t)kStages * kTileRows * DPAD * 8 + (size_t)kStages * kTileRows * 8 + kStages * 8 +
         4 * (kScoreThreads / 32) * 8;
}

template <int DPAD, int KERN, bool EXACT, bool WITH_PHI>
int launch_score_inst(b200rbf_ctx* h, const ScoreArgs& a) {
  auto kern = score_kernel<DPAD, KERN, EXACT, WITH_PHI>;
  constexpr size_t smem = score_smem_bytes<DPAD>();
  static DeviceOnce attr_set;
  if (attr_set.pending(h->device)) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.done(h->device);
  }
  if (a.occ_out != nullptr) {  // query only: resident CTAs per SM of this instantiation
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(a.occ_out, kern, kScoreThreads, smem));
    return 0;
  }
  const long long blocks = (a.m + kScoreThreads - 1) / kScoreThreads;
  kern<<<(unsigned)blocks, kScoreThreads, smem, h->stream>>>(a);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

template <int DPAD>
int launch_score_dpad(b200rbf_ctx* h, const ScoreArgs& a, int kernel, bool exact, bool with_phi) {
  if (!with_phi) {
    return exact ? launch_score_inst<DPAD, 0, true, false>(h, a) : launch_score_inst<DPAD, 0, false, false>(h, a);
  }
  switch (kernel) {
    case B200RBF_KERNEL_CUBIC:
      return exact ? launch_score_inst<DPAD, B200RBF_KERNEL_CUBIC, true, true>(h, a)
                   : launch_score_inst<DPAD, B200RBF_KERNEL_CUBIC, false, true>(h, a);
    case B200RBF_KERNEL_TPS:
      return exact ? launch_score_inst<DPAD, B200RBF_KERNEL_TPS, true, true>(h, a)
                   : launch_score_inst<DPAD, B200RBF_KERNEL_TPS, false, true>(h, a);
    default:
      return exact ? launch_score_inst<DPAD, B200RBF_KERNEL_LINEAR, true, true>(h, a)
                   : launch_score_inst<DPAD, B200RBF_KERNEL_LINEAR, false, true>(h, a);
  }
}

// one translation unit per group of padded dimensions (built in parallel): score_group<G> covers
// DPAD = 8*G + {2, 4, 6, 8}
int launch_score_group0(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group1(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group2(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group3(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group4(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group5(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group6(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);
int launch_score_group7(b200rbf_ctx*, const ScoreArgs&, int, int, bool, bool);

}  // namespace b200rbf
