// Selection half of weighted_distance_merit (candidate_srbf.py:43-55) on device-resident scores.
//   stats_kernel        : x.min() / x.max() of utils.py:60-61 from per-CTA partials
//   merit_argmin_kernel : merit = w*fhat + (1-w)*(1-dhat); merit[dmerit<dtol]=inf; first argmin   (:46-49)
//   pick_kernel         : fvals[jj] = inf; winner = cand[jj]                                       (:50-51)
//   update_kernel       : dmerit = minimum(dmerit, ||cand - winner||) + new min/max partials       (:54-55)
// All of these are HBM-bound streaming passes over m (x d) doubles.
#include "score_kernels.cuh"

namespace b200rbf {

namespace {

constexpr int kSelThreads = 256;

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// numpy argmin ordering: NaN beats everything (np.argmin returns the first NaN), then value, then lowest index
__device__ __forceinline__ bool better(double av, long long ai, double bv, long long bi) {
  const bool an = isnan(av), bn = isnan(bv);
  if (an || bn) {
    if (an && bn) return ai < bi;
    return an;
  }
  if (av < bv) return true;
  if (av > bv) return false;
  return ai < bi;
}

struct Pair {
  double v;
  long long i;
};

__device__ __forceinline__ Pair warp_argmin(Pair p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, p.v, o);
    long long oi = __shfl_xor_sync(0xffffffffu, p.i, o);
    if (better(ov, oi, p.v, p.i)) {
      p.v = ov;
      p.i = oi;
    }
  }
  return p;
}

__device__ Pair block_argmin(Pair p, Pair* s_pairs) {
  p = warp_argmin(p);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_pairs[warp] = p;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    Pair q = lane < nw ? s_pairs[lane] : Pair{d_inf(), 0x7fffffffffffffffLL};
    q = warp_argmin(q);
    if (lane == 0) s_pairs[0] = q;
  }
  __syncthreads();
  return s_pairs[0];
}

// which: bit0 -> reduce f rows (0,1), bit1 -> reduce d rows (2,3)
__global__ void __launch_bounds__(1024) stats_kernel(const double* __restrict__ parts, int stride, int count,
                                                     int which, double* __restrict__ stats) {
  __shared__ double s[4][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[4] = {d_inf(), -d_inf(), d_inf(), -d_inf()};
  for (int i = tid; i < count; i += blockDim.x) {
    if (which & 1) {
      v[0] = fmin(v[0], parts[i]);
      v[1] = fmax(v[1], parts[stride + i]);
    }
    if (which & 2) {
      v[2] = fmin(v[2], parts[2 * stride + i]);
      v[3] = fmax(v[3], parts[3 * stride + i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v[0] = fmin(v[0], __shfl_xor_sync(0xffffffffu, v[0], o));
    v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], o));
    v[2] = fmin(v[2], __shfl_xor_sync(0xffffffffu, v[2], o));
    v[3] = fmax(v[3], __shfl_xor_sync(0xffffffffu, v[3], o));
  }
  if (lane == 0)
    for (int q = 0; q < 4; ++q) s[q][warp] = v[q];
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    double r[4];
    r[0] = lane < nw ? s[0][lane] : d_inf();
    r[1] = lane < nw ? s[1][lane] : -d_inf();
    r[2] = lane < nw ? s[2][lane] : d_inf();
    r[3] = lane < nw ? s[3][lane] : -d_inf();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r[0] = fmin(r[0], __shfl_xor_sync(0xffffffffu, r[0], o));
      r[1] = fmax(r[1], __shfl_xor_sync(0xffffffffu, r[1], o));
      r[2] = fmin(r[2], __shfl_xor_sync(0xffffffffu, r[2], o));
      r[3] = fmax(r[3], __shfl_xor_sync(0xffffffffu, r[3], o));
    }
    if (lane == 0) {
      if (which & 1) {
        stats[0] = r[0];
        stats[1] = r[1];
      }
      if (which & 2) {
        stats[2] = r[2];
        stats[3] = r[3];
      }
    }
  }
}

// stats: {fmin, fmax, dmin, dmax}, or {fmin, -fmax, dmin, -dmax} when negated != 0 (sharded path)
__global__ void __launch_bounds__(kSelThreads) merit_argmin_kernel(const double* __restrict__ fvals,
                                                                   const double* __restrict__ dmerit, long long m,
                                                                   const double* __restrict__ stats, int negated,
                                                                   double w, double dtol, long long idx_off,
                                                                   Pair* __restrict__ out) {
  __shared__ Pair s_pairs[kSelThreads / 32];
  const double fmn = stats[0], fmx = negated ? -stats[1] : stats[1];
  const double dmn = stats[2], dmx = negated ? -stats[3] : stats[3];
  const bool fflat = (fmx == fmn), dflat = (dmx == dmn);  // utils.py:62-63 -> ones
  const double frange = __dsub_rn(fmx, fmn), drange = __dsub_rn(dmx, dmn);
  const double omw = __dsub_rn(1.0, w);
  Pair best{d_inf(), 0x7fffffffffffffffLL};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const double f = fvals[i], dd = dmerit[i];
    // fvals is rescaled once and THEN poisoned with inf at picked rows (candidate_srbf.py:40,50)
    double fh = isinf(f) ? f : (fflat ? 1.0 : __ddiv_rn(__dsub_rn(f, fmn), frange));
    double dh = dflat ? 1.0 : __ddiv_rn(__dsub_rn(dd, dmn), drange);
    double merit = __dadd_rn(__dmul_rn(w, fh), __dmul_rn(omw, __dsub_rn(1.0, dh)));
    if (dd < dtol) merit = d_inf();
    const long long gi = i + idx_off;
    if (better(merit, gi, best.v, best.i)) {
      best.v = merit;
      best.i = gi;
    }
  }
  best = block_argmin(best, s_pairs);
  if (threadIdx.x == 0) out[blockIdx.x] = best;
}

// single CTA: reduce the per-CTA pairs; optionally apply the pick locally
__global__ void __launch_bounds__(kSelThreads) pick_kernel(const Pair* __restrict__ parts, int nparts,
                                                           const double* __restrict__ cand, int d,
                                                           double* __restrict__ fvals, double* __restrict__ winner,
                                                           int dpad, Pair* __restrict__ result_slot,
                                                           Pair* __restrict__ pair_out) {
  __shared__ Pair s_pairs[kSelThreads / 32];
  Pair best{d_inf(), 0x7fffffffffffffffLL};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    Pair p = parts[i];
    if (better(p.v, p.i, best.v, best.i)) best = p;
  }
  best = block_argmin(best, s_pairs);
  if (pair_out != nullptr && threadIdx.x == 0) *pair_out = best;
  if (result_slot != nullptr) {
    if (threadIdx.x == 0) {
      *result_slot = best;
      fvals[best.i] = d_inf();
    }
    double* row_out = reinterpret_cast<double*>(result_slot + 1);  // the slot is {merit, index, row[d]}
    for (int k = threadIdx.x; k < dpad; k += blockDim.x) {
      const double v = k < d ? cand[best.i * (long long)d + k] : 0.0;
      winner[k] = v;
      if (k < d) row_out[k] = v;
    }
  }
}

// ---- sharded path, device-side min-loc (no host round trip between the picks) ----
// single CTA: this rank's best pair and the candidate row behind it -> best = {merit, int64 index bits, row[d]}
__global__ void __launch_bounds__(kSelThreads) best_kernel(const Pair* __restrict__ parts, int nparts,
                                                           const double* __restrict__ cand, int d, long long idx_off,
                                                           double* __restrict__ best_out) {
  __shared__ Pair s_pairs[kSelThreads / 32];
  Pair best{d_inf(), 0x7fffffffffffffffLL};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    Pair p = parts[i];
    if (better(p.v, p.i, best.v, best.i)) best = p;
  }
  best = block_argmin(best, s_pairs);
  if (threadIdx.x == 0) {
    best_out[0] = best.v;
    best_out[1] = __longlong_as_double(best.i);
  }
  const long long local = best.i - idx_off;
  for (int k = threadIdx.x; k < d; k += blockDim.x) best_out[2 + k] = cand[local * (long long)d + k];
}

// single CTA: the winner among the ranks' {merit, index, row} records (np.argmin order: NaN, value, lowest global
// index); result = the winning record, winner = its row padded to dpad; the owner marks the row taken
// (fvals[jj] = inf, candidate_srbf.py:50).  A rank with no candidates contributes index = LLONG_MAX.
__global__ void __launch_bounds__(kSelThreads) decide_kernel(const double* __restrict__ gathered, int world, int d,
                                                             long long idx_off, long long m, double* __restrict__ fvals,
                                                             double* __restrict__ winner, int dpad,
                                                             double* __restrict__ result) {
  __shared__ int s_win;
  if (threadIdx.x == 0) {
    int bw = 0;
    Pair best{gathered[0], __double_as_longlong(gathered[1])};
    for (int r = 1; r < world; ++r) {
      const double* rec = gathered + (size_t)r * (2 + d);
      const Pair p{rec[0], __double_as_longlong(rec[1])};
      if (better(p.v, p.i, best.v, best.i)) {
        best = p;
        bw = r;
      }
    }
    s_win = bw;
    const long long local = best.i - idx_off;
    if (local >= 0 && local < m) fvals[local] = d_inf();
  }
  __syncthreads();
  const double* rec = gathered + (size_t)s_win * (2 + d);
  for (int k = threadIdx.x; k < 2 + d; k += blockDim.x) result[k] = rec[k];
  for (int k = threadIdx.x; k < dpad; k += blockDim.x) winner[k] = k < d ? rec[2 + k] : 0.0;
}

__global__ void take_kernel(const double* __restrict__ cand, int d, long long idx, double* __restrict__ fvals,
                            double* __restrict__ winner, int dpad) {
  if (threadIdx.x == 0) fvals[idx] = d_inf();
  for (int k = threadIdx.x; k < dpad; k += blockDim.x) winner[k] = k < d ? cand[idx * (long long)d + k] : 0.0;
}

// rows are staged through shared memory with coalesced loads, then one thread walks one row sequentially
// (same summation order as cdist, candidate_srbf.py:54)
constexpr int kUpdRows = 128;

template <bool EXACT>
__global__ void __launch_bounds__(kUpdRows) update_kernel(const double* __restrict__ cand, long long m, int d,
                                                          const double* __restrict__ winner,
                                                          double* __restrict__ dmerit, double* __restrict__ parts,
                                                          int nparts, int rpp /* rows per pass <= kUpdRows */) {
  extern __shared__ double s_rows[];  // rpp x (d | 1) doubles, + d winner
  const int ds = d | 1;               // odd stride: conflict-free column walks
  double* s_win = s_rows + (size_t)rpp * ds;
  __shared__ double s_red[2][kUpdRows / 32];
  const int tid = threadIdx.x;
  for (int k = tid; k < d; k += blockDim.x) s_win[k] = winner[k];
  double dmn = d_inf(), dmx = -d_inf();
  for (long long base = (long long)blockIdx.x * rpp; base < m; base += (long long)gridDim.x * rpp) {
    const long long rows = min((long long)rpp, m - base);
    const long long total = rows * d;
    const double* src = cand + base * d;
    __syncthreads();
    // warp w stages rows w, w + nwarps, ...; lanes run over the coordinates: coalesced, no integer division
    (void)total;
    for (int r = tid >> 5; r < (int)rows; r += blockDim.x >> 5) {
      const double* rp = src + (long long)r * d;
      for (int k = tid & 31; k < d; k += 32) s_rows[r * ds + k] = rp[k];
    }
    __syncthreads();
    if (tid < rows) {
      const double* rp = s_rows + tid * ds;
      double acc = 0.0;
      for (int k = 0; k < d; ++k) {
        const double t = rp[k] - s_win[k];
        acc = EXACT ? __dadd_rn(acc, __dmul_rn(t, t)) : fma(t, t, acc);
      }
      const double nd = fmin(dmerit[base + tid], sqrt(acc));
      dmerit[base + tid] = nd;
      dmn = fmin(dmn, nd);
      dmx = fmax(dmx, nd);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmn = fmin(dmn, __shfl_xor_sync(0xffffffffu, dmn, o));
    dmx = fmax(dmx, __shfl_xor_sync(0xffffffffu, dmx, o));
  }
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = dmn;
    s_red[1][tid >> 5] = dmx;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kUpdRows / 32; ++w) {
      dmn = fmin(dmn, s_red[0][w]);
      dmx = fmax(dmx, s_red[1][w]);
    }
    parts[2 * nparts + blockIdx.x] = dmn;
    parts[3 * nparts + blockIdx.x] = dmx;
  }
}

// The same pass with the rows brought in by TMA bulk copies (round 2).  A block of rows is CONTIGUOUS in memory, so one
// cp.async.bulk per block (~50 KB) into a two-stage ring replaces the per-warp row loads: the first version kept ~6 KB of
// loads in flight per SM and reached 1.0 TB/s, 15 % of the HBM peak; Little's law asks for ~45 KB.  Needs 16-byte
// granularity: even d.  Same arithmetic and summation order as update_kernel.
template <bool EXACT>
__global__ void __launch_bounds__(kUpdRows) update_bulk_kernel(const double* __restrict__ cand, long long m, int d,
                                                               const double* __restrict__ winner,
                                                               double* __restrict__ dmerit, double* __restrict__ parts,
                                                               int nparts, int rpp) {
  extern __shared__ __align__(128) double s_blk[];  // 2 x rpp x d rows, + d winner
  double* s_win = s_blk + 2 * (size_t)rpp * d;
  __shared__ uint64_t s_bar[2];
  __shared__ double s_red[2][kUpdRows / 32];
  const int tid = threadIdx.x;
  const long long nblk = (m + rpp - 1) / rpp;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_barrier_init();
  }
  for (int k = tid; k < d; k += blockDim.x) s_win[k] = winner[k];
  __syncthreads();
  auto issue = [&](long long b, int st) {
    const long long base = b * rpp;
    const long long rows = min((long long)rpp, m - base);
    const uint32_t bytes = (uint32_t)(rows * d * 8);
    mbar_expect_tx(&s_bar[st], bytes);
    tma_bulk_g2s(s_blk + (size_t)st * rpp * d, cand + base * d, bytes, &s_bar[st]);
  };
  if (tid == 0) {
    if ((long long)blockIdx.x < nblk) issue(blockIdx.x, 0);
    if ((long long)blockIdx.x + gridDim.x < nblk) issue((long long)blockIdx.x + gridDim.x, 1);
  }
  double dmn = d_inf(), dmx = -d_inf();
  int it = 0;
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x, ++it) {
    const int st = it & 1;
    const long long base = b * rpp;
    const long long rows = min((long long)rpp, m - base);
    mbar_wait(&s_bar[st], (uint32_t)((it >> 1) & 1));
    if (tid < rows) {
      const double* rp = s_blk + (size_t)st * rpp * d + (size_t)tid * d;
      double acc = 0.0;
      for (int k = 0; k < d; ++k) {
        const double t = rp[k] - s_win[k];
        acc = EXACT ? __dadd_rn(acc, __dmul_rn(t, t)) : fma(t, t, acc);
      }
      const double nd = fmin(dmerit[base + tid], sqrt(acc));
      dmerit[base + tid] = nd;
      dmn = fmin(dmn, nd);
      dmx = fmax(dmx, nd);
    }
    __syncthreads();  // the stage is free again
    const long long nb = b + 2 * (long long)gridDim.x;
    if (tid == 0 && nb < nblk) issue(nb, st);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmn = fmin(dmn, __shfl_xor_sync(0xffffffffu, dmn, o));
    dmx = fmax(dmx, __shfl_xor_sync(0xffffffffu, dmx, o));
  }
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = dmn;
    s_red[1][tid >> 5] = dmx;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kUpdRows / 32; ++w) {
      dmn = fmin(dmn, s_red[0][w]);
      dmx = fmax(dmx, s_red[1][w]);
    }
    parts[2 * nparts + blockIdx.x] = dmn;
    parts[3 * nparts + blockIdx.x] = dmx;
  }
}

}  // namespace

int launch_stats(b200rbf_ctx* h, const double* parts, int stride, int count, int which, double* stats) {
  stats_kernel<<<1, 1024, 0, h->stream>>>(parts, stride, count, which, stats);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int launch_merit_argmin(b200rbf_ctx* h, const double* fvals, const double* dmerit, long long m, const double* stats,
                        int stats_negated, double w, double dtol, long long idx_off, void* pair_parts, int* nblocks) {
  long long want = (m + kSelThreads - 1) / kSelThreads;
  long long cap = (long long)h->sm_count * 8;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  merit_argmin_kernel<<<blocks, kSelThreads, 0, h->stream>>>(fvals, dmerit, m, stats, stats_negated, w, dtol, idx_off,
                                                            reinterpret_cast<Pair*>(pair_parts));
  h->launches++;
  B2_CUDA(cudaGetLastError());
  *nblocks = blocks;
  return 0;
}

int launch_pick(b200rbf_ctx* h, const void* pair_parts, int nparts, const double* cand, int d, double* fvals,
                double* winner, int dpad, void* result_slot, void* pair_out) {
  pick_kernel<<<1, kSelThreads, 0, h->stream>>>(reinterpret_cast<const Pair*>(pair_parts), nparts, cand, d, fvals,
                                                winner, dpad, reinterpret_cast<Pair*>(result_slot),
                                                reinterpret_cast<Pair*>(pair_out));
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int launch_best(b200rbf_ctx* h, const void* pair_parts, int nparts, const double* cand, int d, long long idx_off,
                double* best_out) {
  best_kernel<<<1, kSelThreads, 0, h->stream>>>(reinterpret_cast<const Pair*>(pair_parts), nparts, cand, d, idx_off, best_out);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int launch_decide(b200rbf_ctx* h, const double* gathered, int world, int d, long long idx_off, long long m, double* fvals,
                  double* winner, int dpad, double* result) {
  decide_kernel<<<1, kSelThreads, 0, h->stream>>>(gathered, world, d, idx_off, m, fvals, winner, dpad, result);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int launch_take(b200rbf_ctx* h, const double* cand, int d, long long idx, double* fvals, double* winner, int dpad) {
  take_kernel<<<1, 64, 0, h->stream>>>(cand, d, idx, fvals, winner, dpad);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int launch_update(b200rbf_ctx* h, const double* cand, long long m, int d, const double* winner, double* dmerit,
                  double* parts, int stride, int* count_out, bool exact) {
  if (d % 2 == 0 && (reinterpret_cast<uintptr_t>(cand) & 15) == 0 && m >= 4096) {
    // rows per block: 2 stages x rows x d doubles in ~100 KB (2 CTAs per SM), at most one row per thread
    int rpp = (int)((100 * 1024 / sizeof(double) - d) / (2 * (size_t)d));
    if (rpp > kUpdRows) rpp = kUpdRows;
    if (rpp >= 8) {
      const long long want = (m + rpp - 1) / rpp;
      const long long cap = (long long)h->sm_count * 2;
      int blocks = (int)(want < cap ? want : cap);
      if (blocks > stride) blocks = stride;
      const size_t smem = (2 * (size_t)rpp * d + d) * sizeof(double);
      auto kern = exact ? update_bulk_kernel<true> : update_bulk_kernel<false>;
      B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<blocks, kUpdRows, smem, h->stream>>>(cand, m, d, winner, dmerit, parts, stride, rpp);
      h->launches++;
      B2_CUDA(cudaGetLastError());
      *count_out = blocks;
      return 0;
    }
  }
  // rows per pass: as many as fit in ~96 KB of shared memory (2 CTAs per SM), at most kUpdRows
  int rpp = (int)((96 * 1024 / sizeof(double) - d) / (size_t)(d | 1));
  if (rpp > kUpdRows) rpp = kUpdRows;
  B2_CHECK(rpp >= 1, B200RBF_EINVAL, "dimension %d too large for the distance-update kernel", d);
  long long want = (m + rpp - 1) / rpp;
  long long cap = (long long)h->sm_count * 4;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  if (blocks > stride) blocks = stride;
  const size_t smem = ((size_t)rpp * (d | 1) + d) * sizeof(double);
  auto kern = exact ? update_kernel<true> : update_kernel<false>;
  if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, kUpdRows, smem, h->stream>>>(cand, m, d, winner, dmerit, parts, stride, rpp);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  *count_out = blocks;
  return 0;
}

}  // namespace b200rbf
