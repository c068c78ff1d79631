// K2x: fused candidate x center kernel with the squared distance from the norm expansion
//     r^2 = |u|^2 + |c|^2 - 2 u.c
// where the dot products u.c run on the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) and everything else
// (r^2, sqrt, phi, weight FMA, running min) on the FP64 FMA pipe.  On sm_100 both share one datapath (DESIGN.md
// section 4), so the gain over score_kernel is arithmetic, not an extra pipe: d FMA-equivalents per pair instead
// of d subtractions + d FMAs (1.65x fewer FP64-pipe cycles at d = 50), with 15x fewer shared-memory loads because
// the MMA reuses each center element across 32 candidates.
//
// Accuracy.  The expansion's absolute error in r^2 is ~ eps (|u|^2 + |c|^2); where that is not negligible relative to
// r^2 itself, i.e. for NEAR pairs with r^2 < 2^-10 (|u|^2 + |c|^2), the pair is recomputed with direct differences in
// exactly the arithmetic of score_kernel (rare, divergent slow path), so coincident points still give r = 0 exactly
// and `dmerit < dtol` (candidate_srbf.py:48) is decided on a distance with full relative accuracy.  Elsewhere the
// relative error of r^2 is <= ~1e-12, far inside the 1e-10 parity bar; `exact_dist` mode never uses this kernel.
//
// Work decomposition: one warp owns 32 candidates (4 m8 blocks).  Thread (gid = lane / 4, tig = lane % 4) holds
// the A fragments u[gid + 8 i][4 k + tig] in registers for the whole kernel and receives, per 8-center block, the
// dot products of candidates gid + 8 i with centers 2 tig, 2 tig + 1.  Per-candidate sums / minima are therefore
// spread over the 4 threads of a quad and combined with two shuffles at the very end.  Center rows (stride ds == 4
// mod 16 doubles: conflict-free B-fragment loads), their squared norms and weights arrive by TMA bulk copies.
#pragma once
#include "score_kernels.cuh"

namespace b200rbf {

#ifndef B2_MMA_MB
#define B2_MMA_MB 2
#endif
#ifndef B2_MMA_MINB
#define B2_MMA_MINB 4
#endif
constexpr int kMmaThreads = 128;           // 4 warps
constexpr int kMmaMB = B2_MMA_MB;          // m8 candidate blocks per warp (A fragments held in registers)
#ifndef B2_MMA_JB
#define B2_MMA_JB 4
#endif
constexpr int kMmaJB = B2_MMA_JB;          // n8 center blocks per step (8 * JB centers = one tile)
constexpr int kMmaMinBlocks = B2_MMA_MINB; // CTAs per SM the register budget is capped for (one less when d >= 45).
                                           // tools/k2x_tune.cu with the current epilogue (ms per 2^20 x 20000, 3 / 4 / 5
                                           // CTAs): d = 10 - / 28.9 / 30.5, d = 20 42.2 / 41.8 / -, d = 30 52.4 / 52.7 / -,
                                           // d = 50 76.6 / 77.6 / 84.8 (3 CTAs: 168 registers, no spills)
constexpr int kMmaTileRows = 8 * kMmaJB;   // center rows per tile
static_assert(B2_PAD_ROWS % kMmaTileRows == 0, "pad_rows() pads the center rows to multiples of B2_PAD_ROWS: tiles must divide that");
#ifndef B2_MMA_STAGES
#define B2_MMA_STAGES 3
#endif
constexpr int kMmaStages = B2_MMA_STAGES;
constexpr int kMmaRowsPerCta = (kMmaThreads / 32) * 8 * kMmaMB;
constexpr int kMmaPartDoubles = 2 * (kMmaThreads / 32) * 8 * kMmaMB;  // per-warp partial sums and minima (center split)

struct MmaArgs {
  ScoreArgs s;          // pts = shifted centers (npad x ds), coef = weights, as in score_kernel
  const double* cnorm;  // npad squared norms of the shifted centers
  bool aug;             // centers carry (1, |c|^2) in slots d, d + 1 (see score_mma_kernel)
  const double2* tps_table;  // kTpsLogEntries x (inv_j, -0.5 log inv_j), device memory (thin-plate spline only)
  int ds;               // center row stride in doubles (== 4 mod 16, >= 4 * K4)
  int split;            // 1, 2 or 4 warps share one 16-row group and take every split-th center tile (few candidates)
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- thin-plate spline: phi = r^2 log r = c * (0.5 log c), c = r^2.  The 0.5 log c comes from a 128-entry table
// instead of log() (~30 dependent FP64 instructions): c = 2^e m with m in [sqrt(1/2), sqrt(2)); the top 7 fraction bits
// pick a centre m_j with a short reciprocal inv_j (10 significant bits); t = m inv_j - 1 (one FMA, |t| <= 2^-7.9);
//   0.5 log c = e (0.5 ln 2) + (-0.5 log inv_j) + 0.5 log1p(t),   log1p by a degree-7 polynomial (truncation 2e-18).
// 10 FP64 operations, two table words fetched with one LDS.128.  tools/log_probe.cu measures it against 0.5 * log().
constexpr int kTpsLogEntries = 128;

// host: table[j] = (inv_j, -0.5 log inv_j) for fraction index j; j >= kTpsLogSplit belongs to m / 2 in [sqrt(1/2), 1)
constexpr int kTpsLogSplit = 53;  // fraction index of 1.4140625: from there on m is halved and e incremented
inline void tps_log_table_host(double* out /* 2 * kTpsLogEntries */) {
  for (int j = 0; j < kTpsLogEntries; ++j) {
    double mj = 1.0 + (j + 0.5) / kTpsLogEntries;
    if (j >= kTpsLogSplit) mj *= 0.5;
    const double inv = ldexp(nearbyint(ldexp(1.0 / mj, 10)), -10);  // 10 fraction bits: m * inv is exact in the FMA
    out[2 * j] = inv;
    out[2 * j + 1] = (double)(-0.5L * logl((long double)inv));
  }
}

__device__ __forceinline__ double tps_half_log(double c, const double2* __restrict__ tab) {
  const int hi = __double2hiint(c);
  const int j = (hi >> 13) & (kTpsLogEntries - 1);
  const int up = j >= kTpsLogSplit;                                   // m in [sqrt 2, 2): use m / 2, e + 1
  const int e = (hi >> 20) - 1023 + up;
  const double m = __hiloint2double((hi & 0x000fffff) | (up ? 0x3fe00000 : 0x3ff00000), __double2loint(c));
  const double2 tj = tab[j];
  const double t = fma(m, tj.x, -1.0);
  // 0.5 log1p(t) = t (1/2 - t/4 + t^2/6 - t^3/8 + t^4/10 - t^5/12 + t^6/14)
  double q = fma(t, 1.0 / 14.0, -1.0 / 12.0);
  q = fma(t, q, 1.0 / 10.0);
  q = fma(t, q, -1.0 / 8.0);
  q = fma(t, q, 1.0 / 6.0);
  q = fma(t, q, -1.0 / 4.0);
  q = fma(t, q, 0.5);
  // e as a double without the (slow) integer conversion: 2^52 + 2^31 + e has e in its low word
  const double ed = __hiloint2double(0x43300000, (int)(0x80000000u + (unsigned)e)) - 4503601774854144.0;
  return fma(ed, 0.34657359027997264, fma(t, q, tj.y));
}

#ifndef B2_MMA_FAST_SQRT
#define B2_MMA_FAST_SQRT 2
#endif
// phi(r) from r^2 in the DMMA kernel.  The cubic / linear kernels take the square root as B2_MMA_FAST_SQRT corrections
// s <- s + (x - s^2) y / 2 of the hardware reciprocal-square-root seed y (MUFU.RSQ64H, measured 2^-20), starting from
// s = x y.  Measured against sqrt() (tools/sqrt_probe.cu): one correction 1.3e-12, two corrections 2.2e-16 relative
// (within one ulp) -- five FP64 operations, and, unlike sqrt(), no slow-path call, so the 16 chains of a thread sit in
// one basic block and interleave.  0 selects sqrt().
template <int KERN>
__device__ __forceinline__ double mma_phi_from_r2(double r2, const double2* __restrict__ tps_tab) {
  if (KERN == B200RBF_KERNEL_TPS) {
    // tps_kernel.py:28-29 clamps r to eps; here the HIGH WORD of c = r^2 is clamped to that of eps^2 = 2^-104, which
    // keeps log finite at r = 0; below the clamp phi is ~1e-30 either way
    const double c = __hiloint2double(max(__double2hiint(r2), 0x39700000), __double2loint(r2));
    return c * tps_half_log(c, tps_tab);
  }
#if B2_MMA_FAST_SQRT
  if (KERN != B200RBF_KERNEL_TPS) {
    // the seed reads only the high word; clamping it to the smallest normal keeps the seed finite for r2 = 0 (then
    // g = s = 0 exactly) and for subnormal r2 (sqrt < 1e-154: its error is invisible) without any select afterwards
    const int hi = max(__double2hiint(r2), 0x00100000);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(hi, 0)));
    const double g = r2 * y;
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, 0);  // y / 2 on the integer pipe (low word is 0)
    double s = fma(fma(-g, g, r2), h, g);
#if B2_MMA_FAST_SQRT >= 2
    s = fma(fma(-s, s, r2), h, s);  // second correction: ~1 ulp
#endif
    return KERN == B200RBF_KERNEL_CUBIC ? r2 * s : s;
  }
#endif
  return rbf_phi_from_r2<KERN>(r2);
}

// direct squared distance of candidate `row` to the shared-memory center row `cp` (slow path for near pairs)
static __device__ __noinline__ double mma_exact_r2(const double* __restrict__ src, int d, const double* __restrict__ lb,
                                            const double* __restrict__ range, const double* cp) {
  double acc = 0.0;
  for (int k = 0; k < d; ++k) {
    double v = __ldg(src + k);
    if (lb != nullptr) v = __ddiv_rn(__dsub_rn(v, __ldg(lb + k)), __ldg(range + k));
    const double t = v - cp[k];  // same arithmetic as score_kernel: near pairs are bit-identical to the direct path
    acc = fma(t, t, acc);
  }
  return acc;
}

// Near pairs, quad-cooperative (round 2).  The four threads of a quad hold one candidate's coordinates between them
// (dimensions 4 k + tig), so they form the direct squared distance to a shared-memory center row together: K4 FMAs each
// and two shuffles, instead of one thread walking all d coordinates from global memory through the unit-box division
// (mma_exact_r2: ~500 instructions per pair).  On optimiser-like clouds -- hundreds of evaluated points within 0.1 of the
// incumbent, exactly what SOP / DYCORS produce late in a run -- most pairs of a proposal are near pairs: predict() of one
// point took 1.2 us PER CENTER there (3.3 ms at n = 2800, tools/replay_probe.py) against 0.05 us on a uniform cloud.
// Same direct-difference arithmetic, coincident points still give r = 0 exactly.  The whole warp must call it (converged).
template <int K4, bool AUG>
__device__ __forceinline__ double quad_exact_r2(const double (&afr_i)[K4], int d, int tig, const double* __restrict__ crow) {
  double p = 0.0;
#pragma unroll
  for (int k = 0; k < K4; ++k) {
    const int dim = 4 * k + tig;
    if (dim < d) {
      const double u = AUG ? -0.5 * afr_i[k] : afr_i[k];  // the fragments hold -2 u (exact scaling)
      const double t = u - crow[dim];
      p = fma(t, t, p);
    }
  }
  p += __shfl_xor_sync(0xffffffffu, p, 1);
  p += __shfl_xor_sync(0xffffffffu, p, 2);
  return p;
}

// AUG (usable when d + 2 <= 4 K4, i.e. d mod 4 is 1 or 2): the two unused slots of the last k-step carry the norms --
// A row = (-2 u, |u|^2, 1), B row = (c, 1, |c|^2) -- so the MMA accumulates r^2 itself and pass 1 needs no FP64
// instruction at all; the near test then compares against 2^-10 max(|u|^2, |c|^2) (integer max of the high words).
template <int K4, int KERN, bool AUG>
__global__ void __launch_bounds__(kMmaThreads, (K4 >= 12 ? kMmaMinBlocks - 1 : kMmaMinBlocks)) score_mma_kernel(MmaArgs ma) {
  const ScoreArgs& a = ma.s;
  const int ds = ma.ds;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int trows = kMmaTileRows * ma.split;  // center rows per stage: one 32-row tile per warp of a row group
  double* s_pts = reinterpret_cast<double*>(smem_raw);                    // [stages][rows * ds]
  double* s_wts = s_pts + (size_t)kMmaStages * trows * ds;          // [stages][rows]
  double* s_nrm = s_wts + kMmaStages * trows;                       // [stages][rows]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_nrm + kMmaStages * trows);
  double* s_red = reinterpret_cast<double*>(s_full + kMmaStages);         // 4 x warps
  double* s_part = s_red + 4 * (kMmaThreads / 32);                         // center split: warps x 16 rows x (sum, min)
  // thin-plate spline only: the half-log table (2 KB), 16-byte aligned behind the scratch
  const double2* s_tps = reinterpret_cast<const double2*>(s_part + kMmaPartDoubles + (kMmaStages & 1));
  if (KERN == B200RBF_KERNEL_TPS) {
    double2* dst = const_cast<double2*>(s_tps);
    for (int i = threadIdx.x; i < kTpsLogEntries; i += kMmaThreads) dst[i] = ma.tps_table[i];
  }

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  // Center split (few candidates: a strategy's 100 d per proposal would occupy m / 64 CTAs, a fraction of the SMs, and
  // predict() of one point a single warp): `split` warps share a row group, a stage holds `split` 32-row tiles and warp
  // `sub` of the group takes tile `sub` of every stage; the partial sums / minima are combined in a fixed order at the end.
  const int sub = warp & (ma.split - 1);
  const long long base = ((long long)blockIdx.x * ((kMmaThreads / 32) / ma.split) + warp / ma.split) * (8 * kMmaMB);
  const int ntiles = (a.npad + trows - 1) / trows;
  const bool wlive = base < a.m;  // a warp without a live row only keeps the barriers company

  if (tid == 0) {
    for (int s = 0; s < kMmaStages; ++s) mbar_init(&s_full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % kMmaStages;
    const int r0 = t * trows;
    const int rows = min(trows, a.npad - r0);
    mbar_expect_tx(&s_full[s], (uint32_t)rows * (uint32_t)(ds * 8 + 16));
    tma_bulk_g2s(s_pts + (size_t)s * trows * ds, a.pts + (size_t)r0 * ds, (uint32_t)rows * ds * 8u, &s_full[s]);
    tma_bulk_g2s(s_wts + s * trows, a.coef + r0, rows * 8u, &s_full[s]);
    tma_bulk_g2s(s_nrm + s * trows, ma.cnorm + r0, rows * 8u, &s_full[s]);
  };
  if (tid == 0)
    for (int t = 0; t < kMmaStages && t < ntiles; ++t) issue(t);

  // A fragments: shifted unit-box coordinates of candidates base + gid + 8 i, dimensions 4 k + tig
  // All global loads first, then the unit-box divisions: an IEEE division is a call (slow path), the compiler does not
  // move loads across it, and load -> divide -> load -> divide ... cost one memory latency per coordinate -- 2 K4 of them
  // per thread, ~12 us of the ~17 us a one-stage call took (ncu source view of a one-point predict()).
  double afr[kMmaMB][K4];
  double nx[kMmaMB];
  double lbv[K4], rgv[K4];
#pragma unroll
  for (int k = 0; k < K4; ++k) {
    const int dim = 4 * k + tig;
    const bool ok = dim < a.d && a.lb != nullptr;
    lbv[k] = ok ? __ldg(a.lb + dim) : 0.0;
    rgv[k] = ok ? __ldg(a.range + dim) : 1.0;
  }
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    const long long row = base + gid + 8 * i;
    const double* src = a.cand + (row < a.m ? row : a.m - 1) * (long long)a.d;
    const bool live = row < a.m;  // rows past the end compute on the unit-box origin: copies of row m - 1 would repeat
#pragma unroll                    // its near-pair work 15 times over in a one-point predict()
    for (int k = 0; k < K4; ++k) afr[i][k] = (4 * k + tig < a.d && live) ? __ldg(src + 4 * k + tig) : 0.0;
  }
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    const bool live = base + gid + 8 * i < a.m;
    double n2 = 0.0;
#pragma unroll
    for (int k = 0; k < K4; ++k) {
      const int dim = 4 * k + tig;
      double v = afr[i][k];
      if (dim < a.d && live && a.lb != nullptr) v = __ddiv_rn(__dsub_rn(v, lbv[k]), rgv[k]);
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

  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double sum[kMmaMB];
  long long mnb[kMmaMB];  // running min of r^2 as a bit pattern (non-negative doubles order like integers)
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    sum[i] = 0.0;
    mnb[i] = 0x7ff0000000000000LL;
  }
  // Which pairs are recomputed directly (round 2, after the SOP replay: on an optimiser's clustered cloud MOST pairs have
  // r^2 < 2^-10 (|u|^2 + |c|^2), and recomputing all of them made the exception the rule).  The norm expansion's absolute
  // error in r^2 is E ~ 3 eps (|u|^2 + |c|^2) whatever r is, so for the kernel SUM a close pair is no worse than a far
  // one: d(r^3) = 1.5 r E shrinks with r, d(r^2 log r) = E (log r + 1/2) grows only like |log r| (17 E at r = 1e-7; a
  // coincident pair contributes ~1e-14 w instead of 0).  Direct recomputation is needed for
  //  * tiny pairs, r^2 < 2^-T (|u|^2 + |c|^2), T = 26 (r < 1e-4 |u|): sqrt must not see a negative r^2 (a negative double
  //    is a negative integer, below any threshold), coincident points give r = 0 exactly -- and the minimum test below
  //    needs E / r^2 < 2^-20 for every pair it judges by its expanded r^2: E <= ~K4 eps (|u|^2 + |c|^2) = 2^-49 ... and
  //    2^-49 2^26 = 2^-23.  (T = 40 passed every golden and trace but left pairs with 2^-40 < rho < 2^-31 whose
  //    expanded r^2 can sit above the running minimum while the true one is below it: a 1e-5 relative error in dmerit on
  //    a cloud with 1e-6-wide clusters, tests/test_gpu_parity.py::test_call_sizes_around_the_center_split_...)
  //    The linear kernel keeps T = 10, the old rule: d(r) = E / (2 r) is unbounded;
  //  * near pairs (2^-10) that can lower the running minimum: dmerit must be exact relative to ITSELF.  The test is on
  //    the high words (hi(r2) <= hi(min)): 2^-20 of slack, which covers E / r^2 for every non-tiny pair.
  // A near pair that is neither keeps its expanded r^2: it cannot be the minimum and its kernel value is accurate.
  constexpr int kNearShiftHi = 10 << 20;
  constexpr int kTinyShiftHi = (KERN == B200RBF_KERNEL_LINEAR ? 10 : 26) << 20;

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kMmaStages;
    const int rows = min(trows, a.npad - t * trows);  // always a full tile (see pad_rows)
    mbar_wait(&s_full[s], (uint32_t)((t / kMmaStages) & 1));
    const double* tp = s_pts + (size_t)s * trows * ds;
    const double* tw = s_wts + s * trows;
    const double* tn = s_nrm + s * trows;
    const int j0 = 8 * kMmaJB * sub;  // this warp's tile of the stage (npad is a multiple of 32 = 8 kMmaJB: tiles are full)
    if (wlive && j0 < rows) {
      double acc[kMmaMB][kMmaJB][2];
#pragma unroll
      for (int i = 0; i < kMmaMB; ++i)
#pragma unroll
        for (int jb = 0; jb < kMmaJB; ++jb) acc[i][jb][0] = acc[i][jb][1] = 0.0;
      const double* pb = tp + (j0 + gid) * ds + tig;
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
      // Epilogue in three passes so that the common case is branch-free and the 16 sqrt chains of a thread can be
      // interleaved by the scheduler (a call inside the per-pair sequence serialises them):
      // (1) r^2 = |u|^2 + |c|^2 - 2 u.c in place of the dot products, near flags OR-ed together.
      // The near test r2 < 2^-10 sq runs on the INTEGER pipe (the FP64 pipe is the bottleneck) and on the high words
      // only: for non-negative doubles the bit patterns order like the values, 2^-10 sq is sq with its exponent
      // lowered by 10, and a negative r2 (cancellation) is a negative integer, i.e. below any threshold.
      // One flag per pair, two integer instructions: hi(r2) < max(hi(thr) - T, hi(min) + 1), i.e. "tiny, or may lower the
      // running minimum" -- both rare (a row's minimum is lowered O(log n) times over a sweep), so the block below is
      // cold on uniform and clustered clouds alike.
      bool flag = false;
      int pmn[kMmaMB];  // max(hi(|u|^2) - T, hi(min) + 1): the per-row half of the threshold
#pragma unroll
      for (int i = 0; i < kMmaMB; ++i) {
        const int m1 = (int)(mnb[i] >> 32) + 1;
        pmn[i] = AUG ? max(__double2hiint(nx[i]) - kTinyShiftHi, m1) : m1;
      }
#pragma unroll
      for (int jb = 0; jb < kMmaJB; ++jb) {
        const double2 n2 = *reinterpret_cast<const double2*>(tn + j0 + 8 * jb + 2 * tig);
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (AUG) {
              flag |= __double2hiint(acc[i][jb][e]) < max(__double2hiint(e ? n2.y : n2.x) - kTinyShiftHi, pmn[i]);
            } else {
              const double sq = nx[i] + (e ? n2.y : n2.x);
              acc[i][jb][e] = fma(-2.0, acc[i][jb][e], sq);
              flag |= __double2hiint(acc[i][jb][e]) < max(__double2hiint(sq) - kTinyShiftHi, pmn[i]);
            }
          }
        }
      }
      // (2) cold, warp-uniform control flow (every shuffle below runs converged): direct recomputation of the flagged
      // pairs that are tiny or near -- the quad works together on each (quad_exact_r2) -- then the exact update of the
      // running minimum, shared across the quad (four times fewer "new minimum" events than per-thread minima, and
      // with them four times fewer entries into this block: ~12 % of the steps at n = 20 000)
      if (__any_sync(0xffffffffu, flag)) {
        bool rec = false;
#pragma unroll
        for (int jb = 0; jb < kMmaJB; ++jb) {
          const double2 n2 = *reinterpret_cast<const double2*>(tn + j0 + 8 * jb + 2 * tig);
#pragma unroll
          for (int i = 0; i < kMmaMB; ++i) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const double cn = e ? n2.y : n2.x;
              const int thr = AUG ? max(__double2hiint(nx[i]), __double2hiint(cn)) : __double2hiint(nx[i] + cn);
              rec |= __double2hiint(acc[i][jb][e]) < thr - kNearShiftHi;  // superset of the test below, cheaper
            }
          }
        }
        if (__any_sync(0xffffffffu, rec)) {
#pragma unroll
          for (int jb = 0; jb < kMmaJB; ++jb) {
            const double2 n2 = *reinterpret_cast<const double2*>(tn + j0 + 8 * jb + 2 * tig);
#pragma unroll
            for (int i = 0; i < kMmaMB; ++i) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const double cn = e ? n2.y : n2.x;
                const int thr = AUG ? max(__double2hiint(nx[i]), __double2hiint(cn)) : __double2hiint(nx[i] + cn);
                const bool mine = __double2hiint(acc[i][jb][e]) <
                                  max(thr - kTinyShiftHi, min(thr - kNearShiftHi, (int)(mnb[i] >> 32) + 1));
                const unsigned bal = __ballot_sync(0xffffffffu, mine);
                if (bal != 0u) {
#pragma unroll 1
                  for (int src = 0; src < 4; ++src) {
                    if ((bal & (0x11111111u << src)) == 0u) continue;  // no quad has this column flagged
                    const double r2x = quad_exact_r2<K4, AUG>(afr[i], a.d, tig, tp + (j0 + 8 * jb + 2 * src + e) * ds);
                    if (src == tig && mine) acc[i][jb][e] = r2x;
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i) {
#pragma unroll
          for (int jb = 0; jb < kMmaJB; ++jb)
#pragma unroll
            for (int e = 0; e < 2; ++e) mnb[i] = min(mnb[i], __double_as_longlong(acc[i][jb][e]));
          mnb[i] = min(mnb[i], __shfl_xor_sync(0xffffffffu, mnb[i], 1));
          mnb[i] = min(mnb[i], __shfl_xor_sync(0xffffffffu, mnb[i], 2));
        }
      }
      // (3) the weighted kernel sum, same summation order as before
#pragma unroll
      for (int jb = 0; jb < kMmaJB; ++jb) {
        const double2 w2 = *reinterpret_cast<const double2*>(tw + j0 + 8 * jb + 2 * tig);
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i) {
#pragma unroll
          for (int e = 0; e < 2; ++e) sum[i] = fma(mma_phi_from_r2<KERN>(acc[i][jb][e], s_tps), e ? w2.y : w2.x, sum[i]);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && t + kMmaStages < ntiles) issue(t + kMmaStages);
  }

  // combine the quad's partial results (fixed order: deterministic)
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
    mnb[i] = min(mnb[i], __shfl_xor_sync(0xffffffffu, mnb[i], 1));
    mnb[i] = min(mnb[i], __shfl_xor_sync(0xffffffffu, mnb[i], 2));
  }
  if (ma.split > 1) {  // combine the warps of a row group: sub 0 adds the others' partial results, in order
    long long* s_pmin = reinterpret_cast<long long*>(s_part + kMmaPartDoubles / 2);
    if (tig == 0) {
#pragma unroll
      for (int i = 0; i < kMmaMB; ++i) {
        s_part[warp * (8 * kMmaMB) + gid + 8 * i] = sum[i];
        s_pmin[warp * (8 * kMmaMB) + gid + 8 * i] = mnb[i];
      }
    }
    __syncthreads();
    if (sub == 0) {
      for (int o = 1; o < ma.split; ++o) {
#pragma unroll
        for (int i = 0; i < kMmaMB; ++i) {
          sum[i] += s_part[(warp + o) * (8 * kMmaMB) + gid + 8 * i];
          mnb[i] = min(mnb[i], s_pmin[(warp + o) * (8 * kMmaMB) + gid + 8 * i]);
        }
      }
    }
  }
  // tail: lambda_0 + sum_k lambda_k u_k over this thread's dimensions, then across the quad
  double fmn = inf, fmx = -inf, dmn = inf, dmx = -inf;
#pragma unroll
  for (int i = 0; i < kMmaMB; ++i) {
    double ts = 0.0;
    if (a.tail == B200RBF_TAIL_LINEAR) {
#pragma unroll
      for (int k = 0; k < K4; ++k) {
        const int dim = 4 * k + tig;
        if (dim < a.d) ts = fma(__ldg(a.tailc + 1 + dim), afr[i][k], ts);
      }
    }
    ts += __shfl_xor_sync(0xffffffffu, ts, 1);
    ts += __shfl_xor_sync(0xffffffffu, ts, 2);
    if (AUG) ts *= -0.5;  // the fragments hold -2 u (exact scaling)
    const long long row = base + gid + 8 * i;
    const bool live = row < a.m && sub == 0;
    const double fval = sum[i] + (ts + __ldg(a.tailc));
    double dval = 0.0;
    if (a.min_mode != 0) dval = a.dscale * sqrt(__longlong_as_double(mnb[i]));
    if (live && tig == 0) {
      a.fvals[row] = fval;
      if (a.min_mode != 0) {
        if (a.min_mode == 2) dval = fmin(dval, a.dmerit[row]);
        a.dmerit[row] = dval;
      }
    }
    if (a.min_mode == 2) dval = __shfl_sync(0xffffffffu, dval, lane & ~3);  // the combined value lives in tig 0
    if (live) {
      fmn = fmin(fmn, fval);
      fmx = fmax(fmx, fval);
      dmn = fmin(dmn, dval);
      dmx = fmax(dmx, dval);
    }
  }
  fmn = warp_min(fmn);
  fmx = warp_max(fmx);
  dmn = warp_min(dmn);
  dmx = warp_max(dmx);
  constexpr int NW = kMmaThreads / 32;
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
    a.parts[slot] = fmn;
    a.parts[a.nparts + slot] = fmx;
    if (a.min_mode != 0) {
      a.parts[2 * a.nparts + slot] = dmn;
      a.parts[3 * a.nparts + slot] = dmx;
    }
  }
}

inline size_t score_mma_smem_bytes(int ds, bool tps = false, int split = 1) {
  const size_t trows = (size_t)kMmaTileRows * split;
  return ((size_t)kMmaStages * trows * ds + 2 * (size_t)kMmaStages * trows + kMmaStages +
          4 * (kMmaThreads / 32) + kMmaPartDoubles + (tps ? 2 * kTpsLogEntries + 1 : 0)) * 8;
}

template <int K4, int KERN, bool AUG>
int launch_score_mma_inst(b200rbf_ctx* h, const MmaArgs& ma) {
  auto kern = score_mma_kernel<K4, KERN, AUG>;
  const size_t smem = score_mma_smem_bytes(ma.ds, KERN == B200RBF_KERNEL_TPS, ma.split);
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ma.s.occ_out != nullptr) {
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ma.s.occ_out, kern, kMmaThreads, smem));
    return 0;
  }
  const int rows_per_cta = kMmaRowsPerCta / ma.split;
  const long long blocks = (ma.s.m + rows_per_cta - 1) / rows_per_cta;
  kern<<<(unsigned)blocks, kMmaThreads, smem, h->stream>>>(ma);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

template <int K4>
int launch_score_mma_k4(b200rbf_ctx* h, const MmaArgs& ma, int kernel) {
  switch (kernel) {
    case B200RBF_KERNEL_CUBIC:
      return ma.aug ? launch_score_mma_inst<K4, B200RBF_KERNEL_CUBIC, true>(h, ma)
                    : launch_score_mma_inst<K4, B200RBF_KERNEL_CUBIC, false>(h, ma);
    case B200RBF_KERNEL_TPS:
      return ma.aug ? launch_score_mma_inst<K4, B200RBF_KERNEL_TPS, true>(h, ma)
                    : launch_score_mma_inst<K4, B200RBF_KERNEL_TPS, false>(h, ma);
    default:
      return ma.aug ? launch_score_mma_inst<K4, B200RBF_KERNEL_LINEAR, true>(h, ma)
                    : launch_score_mma_inst<K4, B200RBF_KERNEL_LINEAR, false>(h, ma);
  }
}

// K4 = 1..16 (d <= 64), four per translation unit
int launch_score_mma_group0(b200rbf_ctx*, const MmaArgs&, int k4, int kernel);
int launch_score_mma_group1(b200rbf_ctx*, const MmaArgs&, int k4, int kernel);
int launch_score_mma_group2(b200rbf_ctx*, const MmaArgs&, int k4, int kernel);
int launch_score_mma_group3(b200rbf_ctx*, const MmaArgs&, int k4, int kernel);

}  // namespace b200rbf
