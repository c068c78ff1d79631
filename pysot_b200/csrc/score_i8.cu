// K2i: fused candidate x center kernel with the dot products of the norm expansion on the INT8 tensor cores
// (tcgen05.mma.kind::i8, accumulators in TMEM) -- an Ozaki-style splitting made exact by the problem's geometry.
//
// Unit-box coordinates lie in [0, 1], so they are FIXED-point numbers: U = round(u 2^46) = sum_i a_i 256^(5-i) with six
// balanced base-256 digits a_i in [-128, 127] (the top digit is <= 65).  For quantised candidates and centers
//     u.c = 2^-92 sum_{i,j} (a_i . b_j) 256^(10-i-j) = 2^-92 sum_w S_w 256^(10-w),     S_w = sum_{i+j=w} a_i . b_j
// and every S_w is an EXACT int32 tensor-core result: the digit pairs of equal weight w are chained into the same TMEM
// accumulator (|S_w| <= 6 d 2^14 < 2^31), i.e. the pre-summation by weight costs nothing.  Weights w <= 6 are kept (26
// of the 36 digit pairs, seven accumulators = 448 of the 512 TMEM columns): the dropped terms are below 2^-54 d (worst
// case), the quantisation itself moves a coordinate by <= 2^-47 -- together a smaller error in r^2 than the FP64 norm
// expansion's eps (|u|^2 + |c|^2).  (With weights w <= 5 only, the first version, the error was ~40x larger: 6.7e-11 on the
// bench-shape golden against 6.2e-12 for the DMMA kernel.)  Then  r^2 = |u~|^2 + |c~|^2 - 2^-59 V,  V = sum_w S_w 256^(6-w)
// (integer pipe + two additions + one FMA), and the rest -- near-pair recomputation, square root / half-log, weighted
// sum, running minimum, statistics -- is the epilogue of score_mma_kernel.  Per pair the FP64 pipe now executes ~12
// operations instead of 2 x 52 + 7 (d = 50).
//
// Layout.  One CTA = 128 candidates (M) against all centers in tiles of 64 (N).  Operands are K-major int8 rows of 128 B
// in the 128-byte swizzle (UMMA descriptor: SBO = 1024 B, layout SWIZZLE_128B; validated in isolation by
// tools/umma_probe.cu): the candidates' six digit planes are built in shared memory once per CTA (unit-box map,
// quantisation and digit extraction fused), the centers' planes come from a pre-sliced copy made when the model is
// installed, tile by tile through a two-stage cp.async ring.  One elected thread issues the 26 NK MMAs (NK = ceil(d / 32)
// k-steps of 32 bytes) of a tile and commits them to an mbarrier; the sixteen warps then drain their TMEM lane quarter
// (tcgen05.ld 32x32b) -- warp w: candidates 32 (w & 3) .., centers 16 (w >> 2) .. of the tile -- into 16 doubles per
// thread, after which the NEXT tile's MMAs are issued and run underneath the FP64 part of the epilogue.
// Candidates outside [0, 1] (possible through predict()) raise a flag and the caller re-runs the call on the DMMA kernel.
#include "score_mma.cuh"

namespace b200rbf {

constexpr int kI8Rows = 128;      // candidates per CTA (UMMA M)
constexpr int kI8Cols = 64;       // centers per tile (UMMA N)
constexpr int kI8Slices = 6;      // base-256 digits kept per coordinate
constexpr int kI8Weights = 7;     // accumulators: digit pairs of weight i + j = 0 .. 6 (26 of the 36 pairs)
constexpr int kI8EpiThreads = 512;                   // 16 warps: every warp takes part in the epilogue,
constexpr int kI8Threads = kI8EpiThreads;            // thread 0 also issues the MMAs
constexpr int kI8Groups = kI8EpiThreads / 128;       // column groups of the epilogue (warps sharing a TMEM lane quarter)
constexpr int kI8GroupCols = kI8Cols / kI8Groups;    // centers per thread and tile
constexpr int kI8RowBytes = 128;  // one swizzle atom: up to 128 coordinates
constexpr double kI8Scale = 70368744177664.0;            // 2^46
constexpr double kI8InvScale = 1.0 / 70368744177664.0;   // 2^-46
constexpr uint32_t kI8TmemCols = 512;                    // 7 x 64 accumulator columns, rounded to a power of two

struct I8Args {
  ScoreArgs s;              // cand / lb / range / coef / tailc / fvals / dmerit / parts as in score_kernel; pts = fp64 centers
  const int8_t* cq;         // sliced centers: [tile][slice][64 rows][32 nk bytes]
  const double* cnq;        // squared norms of the quantised centers (padded to tiles)
  int dpad;                 // row stride of s.pts (fp64 centers, for the near-pair recomputation)
  int nk, ntiles;
  int* flag;                // set when a candidate coordinate falls outside [0, 1]
  const double2* tps_table;
};

__host__ __device__ inline int i8_sw128(int r, int k) {
  return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 4) ^ (r & 7)) & 7) << 4) + (k & 15);
}

// six balanced base-256 digits of U = round(u 2^46), most significant first; uq = U 2^-46
__device__ __forceinline__ bool i8_digits(double u, int dg[kI8Slices], double* uq) {
  const bool ok = u >= 0.0 && u <= 1.0;
  long long U = __double2ll_rn((ok ? u : 0.0) * kI8Scale);
  *uq = (double)U * kI8InvScale;
#pragma unroll
  for (int i = kI8Slices - 1; i >= 0; --i) {
    const int dgt = (int)((U + 128) & 255) - 128;
    dg[i] = dgt;
    U = (U - dgt) >> 8;
  }
  return ok;
}

// pre-sliced centers + norms of the quantised centers; one thread per padded row (pad rows repeat the last center)
__global__ void i8_slice_centers_kernel(const double* __restrict__ Xu, int stride, int n, int d, int nrows, int nk,
                                        int8_t* __restrict__ cq, double* __restrict__ cnq, int* __restrict__ flag) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const double* src = Xu + (size_t)(row < n ? row : n - 1) * stride;
  const int tile = row / kI8Cols, r = row % kI8Cols;
  const int kb = 32 * nk;
  int8_t* base = cq + (size_t)tile * kI8Slices * kI8Cols * kb + (size_t)r * kb;
  double n2 = 0.0;
  bool ok = true;
  for (int k = 0; k < kb; ++k) {
    int dg[kI8Slices] = {0, 0, 0, 0, 0, 0};
    if (k < d) {
      double uq;
      ok &= i8_digits(src[k], dg, &uq);
      n2 = fma(uq, uq, n2);
    }
#pragma unroll
    for (int i = 0; i < kI8Slices; ++i) base[(size_t)i * kI8Cols * kb + k] = (int8_t)dg[i];
  }
  cnq[row] = n2;
  if (!ok) atomicExch(flag, 1);
}

__device__ __forceinline__ uint64_t i8_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void i8_mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void i8_tmem_ld8(uint32_t addr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr));
}
__device__ __forceinline__ void i8_tmem_ld16(uint32_t addr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
}
// int64 (|x| < 2^51) -> double.  B2_I8_XU_CONV = 1: the conversion instruction (XU pipe, quarter rate; measured: no faster)
// apart from one MUFU per pair); 0: the 2^52 + 2^51 bias trick (two integer adds + one FP64 add)
#ifndef B2_I8_XU_CONV
#define B2_I8_XU_CONV 0
#endif
#if B2_I8_XU_CONV
#define I8_CONV(x) ((double)(x))
#else
#define I8_CONV(x) (__longlong_as_double((x) + 0x4338000000000000LL) - 6755399441055744.0)
#endif
__device__ __forceinline__ void cp_async16_i8(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

template <int NK, int KERN>
__global__ void __launch_bounds__(kI8Threads, 1) score_i8_kernel(I8Args ia) {
  const ScoreArgs& a = ia.s;
  extern __shared__ unsigned char smem_i8[];
  // the 128-byte swizzle is a function of the shared-memory ADDRESS bits: operand tiles must start 1024-byte aligned
  unsigned char* sA = smem_i8 + ((1024u - (smem_u32(smem_i8) & 1023u)) & 1023u);  // 6 x 128 x 128 B
  unsigned char* sB = sA + kI8Slices * kI8Rows * kI8RowBytes;                // 2 stages x 6 x 64 x 128 B
  double* sW = reinterpret_cast<double*>(sB + 2 * kI8Slices * kI8Cols * kI8RowBytes);  // 4 x 64 weights
  double* sN = sW + 4 * kI8Cols;                                             // 4 x 64 center norms
  double* sNu = sN + 4 * kI8Cols;                                            // 128 candidate norms
  double* sTail = sNu + kI8Rows;                                             // 128 tail values
  double* sSum = sTail + kI8Rows;                                            // groups x 128 partial sums
  long long* sMin = reinterpret_cast<long long*>(sSum + kI8Groups * kI8Rows);  // groups x 128 partial minima
  double* sRed = reinterpret_cast<double*>(sMin + kI8Groups * kI8Rows);      // 4 x 16 statistics
  const double2* s_tps = reinterpret_cast<const double2*>(sRed + 64);        // half-log table (TPS)
  __shared__ uint64_t mbar;        // "the MMAs of the low-order accumulators (weights 3 .. 6) of tile t have completed"
  __shared__ uint64_t mbar_free;   // "... and those of the high-order accumulators (weights 0 .. 2)"  (tcgen05.commit)
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)blockIdx.x * kI8Rows;
  constexpr int KB = 32 * NK;  // bytes (= coordinates) per operand row actually used

  if (KERN == B200RBF_KERNEL_TPS)
    for (int i = tid; i < kTpsLogEntries; i += kI8Threads) const_cast<double2*>(s_tps)[i] = ia.tps_table[i];
  // zero the candidate planes (padding coordinates must contribute nothing)
  for (int e = tid; e < kI8Slices * kI8Rows * kI8RowBytes / 16; e += kI8Threads) reinterpret_cast<uint4*>(sA)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    mbar_init(&mbar_free, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(kI8TmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // center tile loader: 16-byte chunks of the compact pre-sliced rows into the swizzled 128-byte rows (2-stage ring);
  // weights and norms of the tile's centers go to a 4-stage ring (they are still read while tile t + 2 is in flight)
  auto load_tile = [&](int t) {
    const int st = t & 1, sw = t & 3;
    unsigned char* dstB = sB + (size_t)st * kI8Slices * kI8Cols * kI8RowBytes;
    const int8_t* src = ia.cq + (size_t)t * kI8Slices * kI8Cols * KB;
    constexpr int CH = KB / 16;  // chunks per row
    for (int q = tid; q < kI8Slices * kI8Cols * CH; q += kI8EpiThreads) {
      const int sl = q / (kI8Cols * CH), rem = q - sl * (kI8Cols * CH), r = rem / CH, ch = rem - r * CH;
      cp_async16_i8(dstB + sl * (kI8Cols * kI8RowBytes) + (r >> 3) * 1024 + (r & 7) * 128 + (((ch ^ (r & 7)) & 7) << 4),
                    src + (size_t)q * 16);
    }
    if (tid < kI8Cols / 2) {
      cp_async16_i8(sW + sw * kI8Cols + 2 * tid, a.coef + (size_t)t * kI8Cols + 2 * tid);
    } else if (tid < kI8Cols) {
      const int j = tid - kI8Cols / 2;
      cp_async16_i8(sN + sw * kI8Cols + 2 * j, ia.cnq + (size_t)t * kI8Cols + 2 * j);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_tile(0);
  __syncthreads();  // planes zeroed before the digits are written

  // candidate digit planes: thread = (row, quarter of the coordinates)
  {
    const int row = tid >> 2, part = tid & 3;
    const long long grow = base + row < a.m ? base + row : a.m - 1;
    const double* src = a.cand + grow * (long long)a.d;
    double n2 = 0.0, ts = 0.0;
    bool ok = true;
    for (int k = part; k < a.d; k += 4) {
      double v = __ldg(src + k);
      if (a.lb != nullptr) v = __ddiv_rn(__dsub_rn(v, __ldg(a.lb + k)), __ldg(a.range + k));  // utils.py:33
      if (a.tail == B200RBF_TAIL_LINEAR) ts = fma(__ldg(a.tailc + 1 + k), v, ts);
      int dg[kI8Slices];
      double uq;
      ok &= i8_digits(v, dg, &uq);
      n2 = fma(uq, uq, n2);
      const int off = i8_sw128(row, k);
#pragma unroll
      for (int i = 0; i < kI8Slices; ++i) sA[i * (kI8Rows * kI8RowBytes) + off] = (unsigned char)dg[i];
    }
    n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
    n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
    ts += __shfl_xor_sync(0xffffffffu, ts, 1);
    ts += __shfl_xor_sync(0xffffffffu, ts, 2);
    if (part == 0) {
      sNu[row] = n2;
      sTail[row] = ts + __ldg(a.tailc);
    }
    if (!ok) atomicExch(ia.flag, 1);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kI8Cols >> 3) << 17) | ((uint32_t)(kI8Rows >> 4) << 24);
  const uint64_t descA = i8_desc(smem_u32(sA));
  // the 26 NK MMAs of one tile: accumulator w (columns 64 w ..) receives the digit pairs (i, w - i).  The low-order
  // accumulators (w = 3 .. 6, 20 NK MMAs) go first and are committed on their own: the epilogue drains and combines them
  // on the INTEGER pipe while the high-order ones (w = 0 .. 2, 6 NK MMAs) still execute -- integer work and tcgen05 MMAs
  // overlap; FP64 work and MMAs do not (measured, tools/umma_overlap.cu: a DFMA loop and an MMA stream in the same CTA take
  // the SUM of their times, an integer loop and the MMA stream the maximum).
  auto issue_weights = [&](int t, int w_lo, int w_hi, uint64_t* bar) {
    const uint64_t descB = i8_desc(smem_u32(sB + (size_t)(t & 1) * kI8Slices * kI8Cols * kI8RowBytes));
#pragma unroll
    for (int w = 0; w < kI8Weights; ++w) {
      if (w < w_lo || w > w_hi) continue;
      bool first = true;
#pragma unroll
      for (int i = 0; i < kI8Slices; ++i) {
        if (w - i < 0 || w - i >= kI8Slices) continue;
#pragma unroll
        for (int c = 0; c < NK; ++c) {
          const uint64_t da = descA + (uint64_t)(i * (kI8Rows * kI8RowBytes / 16) + 2 * c);
          const uint64_t db = descB + (uint64_t)((w - i) * (kI8Cols * kI8RowBytes / 16) + 2 * c);
          i8_mma(tmem + (uint32_t)(w * kI8Cols), da, db, idesc, first ? 0u : 1u);
          first = false;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  };
  auto issue_tile = [&](int t) {
    issue_weights(t, 3, 6, &mbar);
    issue_weights(t, 0, 2, &mbar_free);
  };
  // which pairs are recomputed directly: tiny ones and near ones that can lower the running minimum (score_mma.cuh)
  constexpr int kNearShiftHi = 10 << 20;
  constexpr int kTinyShiftHi = (KERN == B200RBF_KERNEL_LINEAR ? 10 : 26) << 20;
  const int q4 = warp & 3, cg = (warp >> 2) & 3;  // TMEM lane quarter = warp id within the CTA mod 4
  const int crow = 32 * q4 + lane;
  double sum = 0.0;
  long long mnb = 0x7ff0000000000000LL;

  if (tid == 0) issue_tile(0);
  if (ia.ntiles > 1) load_tile(1);
  const long long grow = base + crow < a.m ? base + crow : a.m - 1;
  const double nu = sNu[crow];
  uint32_t phase = 0;
  const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(kI8GroupCols * cg);
  auto wait_bar = [&](uint64_t* bar) {  // bounded: a fault in the tensor-core path must surface as an error, not a hung GPU
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, phase)) {
      if (++spins > (1u << 26)) {
        if (tid == 0) atomicExch(ia.flag, 2);
        break;
      }
    }
  };
  for (int t = 0; t < ia.ntiles; ++t) {
    // phase A: drain this thread's 16 x 7 accumulator words into 16 doubles  V = sum_w S_w 256^(6-w), all on the integer
    // pipe except two additions and one FMA: L = S3 2^24 + S4 2^16 + S5 2^8 + S6 and H = S0 2^16 + S1 2^8 + S2 as int64
    // (|L| < 2^46, |H| < 2^34), each made a double by the 2^52 + 2^51 bias trick (the int -> double conversion instruction
    // runs on the quarter-rate XU pipe; ncu showed it as the busiest pipe of the first version).
    double V[kI8GroupCols];
    wait_bar(&mbar);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      uint32_t v[4][kI8GroupCols];
#pragma unroll
      for (int w = 0; w < 4; ++w) i8_tmem_ld16(taddr + (uint32_t)((3 + w) * kI8Cols), v[w]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < kI8GroupCols; ++j) {
        const long long L = (long long)((int)v[0][j] * 256 + (int)v[1][j]) * 65536 + (long long)((int)v[2][j] * 256 + (int)v[3][j]);
        V[j] = I8_CONV(L);
      }
    }
    wait_bar(&mbar_free);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      uint32_t v[3][kI8GroupCols];
#pragma unroll
      for (int w = 0; w < 3; ++w) i8_tmem_ld16(taddr + (uint32_t)(w * kI8Cols), v[w]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < kI8GroupCols; ++j) {
        const long long H = (long long)(int)v[0][j] * 65536 + (long long)((int)v[1][j] * 256 + (int)v[2][j]);
        V[j] = fma(I8_CONV(H), 4294967296.0, V[j]);
      }
    }
    phase ^= 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");          // tile t + 1 has landed ...
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // ... and is visible to the tensor core
    __syncthreads();                                              // every warp has drained its accumulators
    if (t + 1 < ia.ntiles && tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_tile(t + 1);
    }
    if (t + 2 < ia.ntiles) load_tile(t + 2);
    // phase B: r^2 = |u~|^2 + |c~|^2 - 2^-59 V, near-pair recomputation, phi, weighted sum, running minimum
    const double* tw = sW + (t & 3) * kI8Cols + kI8GroupCols * cg;
    const double* tn = sN + (t & 3) * kI8Cols + kI8GroupCols * cg;
#pragma unroll
    for (int c0 = 0; c0 < kI8GroupCols; c0 += 8) {
      double r2v[8];
      // one flag per pair: tiny (direct recomputation) or may lower the running minimum -- both rare (score_mma.cuh)
      bool flag = false;
      const int mnhi1 = (int)(mnb >> 32) + 1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double sq = nu + tn[c0 + j];
        const double r2 = fma(V[c0 + j], -1.734723475976807e-18 /* 2^-59 */, sq);
        flag |= __double2hiint(r2) < max(__double2hiint(sq) - kTinyShiftHi, mnhi1);
        r2v[j] = r2;
      }
      if (flag) {  // rare
        bool near_any = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) near_any |= __double2hiint(r2v[j]) < __double2hiint(nu + tn[c0 + j]) - kNearShiftHi;
        if (near_any) {  // rarer: direct differences (the arithmetic of score_kernel) for the tiny and the near flagged pairs
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int thr = __double2hiint(nu + tn[c0 + j]);
            if (__double2hiint(r2v[j]) < max(thr - kTinyShiftHi, min(thr - kNearShiftHi, mnhi1))) {
              const int cj = t * kI8Cols + kI8GroupCols * cg + c0 + j;
              const int cjr = cj < a.npad ? cj : a.npad - 1;
              r2v[j] = mma_exact_r2(a.cand + grow * (long long)a.d, a.d, a.lb, a.range, a.pts + (size_t)cjr * ia.dpad);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) mnb = min(mnb, __double_as_longlong(r2v[j]));  // pad rows duplicate a real center
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum = fma(mma_phi_from_r2<KERN>(r2v[j], s_tps), tw[c0 + j], sum);
    }
  }
  // combine the column groups, write the outputs and the CTA's statistics
  sSum[cg * kI8Rows + crow] = sum;
  sMin[cg * kI8Rows + crow] = mnb;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double fmn = inf, fmx = -inf, dmn = inf, dmx = -inf;
  if (tid < kI8Rows) {
    const long long row = base + tid;
    if (row < a.m) {
      double fs = 0.0;
      long long mb = 0x7ff0000000000000LL;
#pragma unroll
      for (int g = 0; g < kI8Groups; ++g) {
        fs += sSum[g * kI8Rows + tid];
        mb = min(mb, sMin[g * kI8Rows + tid]);
      }
      const double fval = fs + sTail[tid];
      double dval = 0.0;
      a.fvals[row] = fval;
      if (a.min_mode != 0) {
        dval = a.dscale * sqrt(__longlong_as_double(mb));
        if (a.min_mode == 2) dval = fmin(dval, a.dmerit[row]);
        a.dmerit[row] = dval;
      }
      fmn = fmx = fval;
      dmn = dmx = dval;
    }
  }
  fmn = warp_min(fmn);
  fmx = warp_max(fmx);
  dmn = warp_min(dmn);
  dmx = warp_max(dmx);
  if (lane == 0) {
    sRed[warp] = fmn;
    sRed[16 + warp] = fmx;
    sRed[32 + warp] = dmn;
    sRed[48 + warp] = dmx;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 4; ++w) {  // rows live in warps 0..3
      fmn = fmin(fmn, sRed[w]);
      fmx = fmax(fmx, sRed[16 + w]);
      dmn = fmin(dmn, sRed[32 + w]);
      dmx = fmax(dmx, sRed[48 + w]);
    }
    const int slot = a.part_off + blockIdx.x;
    a.parts[slot] = fmn;
    a.parts[a.nparts + slot] = fmx;
    if (a.min_mode != 0) {
      a.parts[2 * a.nparts + slot] = dmn;
      a.parts[3 * a.nparts + slot] = dmx;
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kI8TmemCols));
}

static size_t i8_smem_bytes(bool tps) {
  return (size_t)kI8Slices * kI8Rows * kI8RowBytes + 2 * (size_t)kI8Slices * kI8Cols * kI8RowBytes +
         (8 * kI8Cols + 2 * kI8Rows + 2 * kI8Groups * kI8Rows + 64) * sizeof(double) + (tps ? 2 * kTpsLogEntries * 8 : 0) + 64 + 1024;
}

template <int NK, int KERN>
static int launch_i8_inst(b200rbf_ctx* h, const I8Args& ia) {
  auto kern = score_i8_kernel<NK, KERN>;
  const size_t smem = i8_smem_bytes(KERN == B200RBF_KERNEL_TPS);
  static DeviceOnce attr;
  if (attr.pending(h->device)) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.done(h->device);
  }
  if (ia.s.occ_out != nullptr) {
    *ia.s.occ_out = 1;
    return 0;
  }
  const long long blocks = (ia.s.m + kI8Rows - 1) / kI8Rows;
  kern<<<(unsigned)blocks, kI8Threads, smem, h->stream>>>(ia);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

template <int NK>
static int launch_i8_nk(b200rbf_ctx* h, const I8Args& ia, int kernel) {
  switch (kernel) {
    case B200RBF_KERNEL_CUBIC: return launch_i8_inst<NK, B200RBF_KERNEL_CUBIC>(h, ia);
    case B200RBF_KERNEL_TPS: return launch_i8_inst<NK, B200RBF_KERNEL_TPS>(h, ia);
    default: return launch_i8_inst<NK, B200RBF_KERNEL_LINEAR>(h, ia);
  }
}

int i8_rows_per_cta() { return kI8Rows; }

// pre-slice the resident centers (lazily, on the first int8 scoring call after a model was installed)
int i8_prepare_model(b200rbf_ctx* h) {
  Model& M = h->model;
  if (M.i8_ready) return 0;
  const int n = M.n, d = M.d;
  M.i8_nk = (d + 31) / 32;
  M.i8_ntiles = (n + kI8Cols - 1) / kI8Cols;
  const int nrows = M.i8_ntiles * kI8Cols;
  B2_TRY(M.centers_q.reserve((size_t)nrows * kI8Slices * 32 * M.i8_nk));
  B2_TRY(M.cnorm_q.reserve((size_t)nrows * sizeof(double)));
  B2_TRY(M.coef_q.reserve((size_t)nrows * sizeof(double)));
  B2_TRY(M.i8_flag.reserve(64));
  B2_CUDA(cudaMemsetAsync(M.i8_flag.p, 0, 8, h->stream));
  B2_CUDA(cudaMemsetAsync(M.coef_q.p, 0, (size_t)nrows * sizeof(double), h->stream));
  B2_CUDA(cudaMemcpyAsync(M.coef_q.p, M.coef.p, (size_t)(M.npad < nrows ? M.npad : nrows) * sizeof(double), cudaMemcpyDeviceToDevice,
                          h->stream));
  i8_slice_centers_kernel<<<(nrows + 127) / 128, 128, 0, h->stream>>>(M.centers.as<double>(), M.dpad, n, d, nrows, M.i8_nk,
                                                                      M.centers_q.as<int8_t>(), M.cnorm_q.as<double>(),
                                                                      M.i8_flag.as<int>());
  h->launches++;
  B2_CUDA(cudaGetLastError());
  M.i8_ready = true;
  return 0;
}

int launch_score_i8(b200rbf_ctx* h, const ScoreArgs& a) {
  if (a.occ_out == nullptr) B2_TRY(i8_prepare_model(h));
  const Model& M = h->model;
  I8Args ia;
  ia.s = a;
  ia.s.pts = M.centers.as<double>();
  ia.s.coef = M.coef_q.as<double>();
  ia.cq = M.centers_q.as<int8_t>();
  ia.cnq = M.cnorm_q.as<double>();
  ia.dpad = M.dpad;
  ia.nk = M.i8_nk;
  ia.ntiles = M.i8_ntiles;
  ia.flag = M.i8_flag.as<int>() + 1;
  ia.tps_table = nullptr;
  if (M.kernel == B200RBF_KERNEL_TPS) {
    int ensure_tps_table(b200rbf_ctx*);
    B2_TRY(ensure_tps_table(h));
    ia.tps_table = h->tps_table.as<double2>();
  }
  switch ((M.d + 31) / 32) {
    case 1: return launch_i8_nk<1>(h, ia, M.kernel);
    case 2: return launch_i8_nk<2>(h, ia, M.kernel);
    case 3: return launch_i8_nk<3>(h, ia, M.kernel);
    case 4: return launch_i8_nk<4>(h, ia, M.kernel);
  }
  set_error("internal: no int8 score kernel for d = %d", M.d);
  return B200RBF_EINVAL;
}

}  // namespace b200rbf
