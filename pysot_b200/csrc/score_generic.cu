// Fused score kernel for padded dimensions above kMaxRegDim (64): same contract as score_kernel, but the
// candidate coordinates live in shared memory ([coordinate][thread], conflict-free) instead of registers and the
// coordinate loop is a runtime loop.  Center tiles still arrive by TMA bulk copy.  Supports dpad <= kGenMaxDim.
#include "score_kernels.cuh"

namespace b200rbf {
namespace {

constexpr int kGenThreads = 128;
constexpr int kGenTileRows = 8;
constexpr int kGenStages = 2;

template <int KERN, bool EXACT, bool WITH_PHI>
__global__ void __launch_bounds__(kGenThreads) score_generic_kernel(ScoreArgs a, int dpad) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_pts = reinterpret_cast<double*>(smem_raw);                     // [stages][rows][dpad]
  double* s_wts = s_pts + (size_t)kGenStages * kGenTileRows * dpad;        // [stages][rows]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_wts + kGenStages * kGenTileRows);
  double* s_red = reinterpret_cast<double*>(s_full + kGenStages);          // 4 x warps
  double* s_x = s_red + 4 * (kGenThreads / 32);                            // [dpad][threads]

  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * kGenThreads + tid;
  const bool live = row < a.m;
  const int ntiles = (a.npad + kGenTileRows - 1) / kGenTileRows;
  const uint32_t row_bytes = (uint32_t)dpad * 8u;

  if (tid == 0) {
    for (int s = 0; s < kGenStages; ++s) mbar_init(&s_full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % kGenStages;
    const int r0 = t * kGenTileRows;
    const int rows = min(kGenTileRows, a.npad - r0);
    mbar_expect_tx(&s_full[s], rows * row_bytes + (WITH_PHI ? rows * 8u : 0u));
    tma_bulk_g2s(s_pts + (size_t)s * kGenTileRows * dpad, a.pts + (size_t)r0 * dpad, rows * row_bytes, &s_full[s]);
    if (WITH_PHI) tma_bulk_g2s(s_wts + s * kGenTileRows, a.coef + r0, rows * 8u, &s_full[s]);
  };
  if (tid == 0)
    for (int t = 0; t < kGenStages && t < ntiles; ++t) issue(t);

  {
    const double* src = a.cand + (live ? row : 0) * (long long)a.d;
    for (int k = 0; k < dpad; ++k) {
      double v = 0.0;
      if (k < a.d) {
        v = __ldg(src + k);
        if (a.lb != nullptr) v = __ddiv_rn(__dsub_rn(v, __ldg(a.lb + k)), __ldg(a.range + k));
      }
      s_x[k * kGenThreads + tid] = v;
    }
  }
  double sum = 0.0;
  double minr2 = __longlong_as_double(0x7ff0000000000000LL);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kGenStages;
    const int rows = min(kGenTileRows, a.npad - t * kGenTileRows);
    mbar_wait(&s_full[s], (uint32_t)((t / kGenStages) & 1));
    const double* tp = s_pts + (size_t)s * kGenTileRows * dpad;
    const double* tw = s_wts + s * kGenTileRows;
    for (int j0 = 0; j0 < rows; j0 += 4) {
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
      const double* c0 = tp + (j0 + 0) * dpad;
      const double* c1 = tp + (j0 + 1) * dpad;
      const double* c2 = tp + (j0 + 2) * dpad;
      const double* c3 = tp + (j0 + 3) * dpad;
#pragma unroll 4
      for (int k = 0; k < dpad; ++k) {
        const double x = s_x[k * kGenThreads + tid];
        const double t0 = x - c0[k], t1 = x - c1[k], t2 = x - c2[k], t3 = x - c3[k];
        if (EXACT) {
          acc0 = __dadd_rn(acc0, __dmul_rn(t0, t0));
          acc1 = __dadd_rn(acc1, __dmul_rn(t1, t1));
          acc2 = __dadd_rn(acc2, __dmul_rn(t2, t2));
          acc3 = __dadd_rn(acc3, __dmul_rn(t3, t3));
        } else {
          acc0 = fma(t0, t0, acc0);
          acc1 = fma(t1, t1, acc1);
          acc2 = fma(t2, t2, acc2);
          acc3 = fma(t3, t3, acc3);
        }
      }
      minr2 = fmin(fmin(minr2, acc0), fmin(acc1, fmin(acc2, acc3)));
      if (WITH_PHI) {
        sum = fma(rbf_phi_from_r2<KERN>(acc0), tw[j0 + 0], sum);
        sum = fma(rbf_phi_from_r2<KERN>(acc1), tw[j0 + 1], sum);
        sum = fma(rbf_phi_from_r2<KERN>(acc2), tw[j0 + 2], sum);
        sum = fma(rbf_phi_from_r2<KERN>(acc3), tw[j0 + 3], sum);
      }
    }
    __syncthreads();
    if (tid == 0 && t + kGenStages < ntiles) issue(t + kGenStages);
  }

  double fval = 0.0, dval = 0.0;
  if (WITH_PHI) {
    double tsum = __ldg(a.tailc);
    if (a.tail == B200RBF_TAIL_LINEAR)
      for (int k = 0; k < a.d; ++k) tsum = fma(__ldg(a.tailc + 1 + k), s_x[k * kGenThreads + tid], tsum);
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
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double fmn = warp_min(live ? fval : inf), fmx = warp_max(live ? fval : -inf);
  double dmn = warp_min(live ? dval : inf), dmx = warp_max(live ? dval : -inf);
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kGenThreads / 32;
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

template <int KERN, bool EXACT, bool WITH_PHI>
int launch_generic_inst(b200rbf_ctx* h, const ScoreArgs& a, int dpad) {
  auto kern = score_generic_kernel<KERN, EXACT, WITH_PHI>;
  const size_t smem = ((size_t)kGenStages * kGenTileRows * dpad + kGenStages * kGenTileRows + kGenStages +
                       4 * (kGenThreads / 32) + (size_t)dpad * kGenThreads) * 8;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (a.occ_out != nullptr) {
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(a.occ_out, kern, kGenThreads, smem));
    return 0;
  }
  const long long blocks = (a.m + kGenThreads - 1) / kGenThreads;
  kern<<<(unsigned)blocks, kGenThreads, smem, h->stream>>>(a, dpad);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int score_generic_threads() { return kGenThreads; }

int launch_score_generic(b200rbf_ctx* h, const ScoreArgs& a, int dpad, int kernel, bool exact, bool with_phi) {
  B2_CHECK(dpad <= kGenMaxDim, B200RBF_EINVAL, "dim (padded %d) > %d is not supported by this build", dpad, kGenMaxDim);
  if (!with_phi)
    return exact ? launch_generic_inst<0, true, false>(h, a, dpad) : launch_generic_inst<0, false, false>(h, a, dpad);
  switch (kernel) {
    case B200RBF_KERNEL_CUBIC:
      return exact ? launch_generic_inst<B200RBF_KERNEL_CUBIC, true, true>(h, a, dpad)
                   : launch_generic_inst<B200RBF_KERNEL_CUBIC, false, true>(h, a, dpad);
    case B200RBF_KERNEL_TPS:
      return exact ? launch_generic_inst<B200RBF_KERNEL_TPS, true, true>(h, a, dpad)
                   : launch_generic_inst<B200RBF_KERNEL_TPS, false, true>(h, a, dpad);
    default:
      return exact ? launch_generic_inst<B200RBF_KERNEL_LINEAR, true, true>(h, a, dpad)
                   : launch_generic_inst<B200RBF_KERNEL_LINEAR, false, true>(h, a, dpad);
  }
}

}  // namespace b200rbf
