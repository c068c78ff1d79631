// Device-side candidate generation: the perturbation half of candidate_dycors / candidate_srbf /
// candidate_uniform without the host.  On the CPU scipy.stats.truncnorm.rvs is ~50 % of candidate_dycors at d = 10 and
// makes BASELINE config #4 host-bound (64 M x 50 candidates take minutes to draw and 25 GB to ship); here a row is one
// warp's worth of Philox rounds and inverse-CDF evaluations.  OPT-IN: the streams differ from numpy's / scipy's, so a seeded run does not reproduce the
// reference's candidate SETS -- the distributions are the same (tests/test_gpu_generate.py: KS tests, mask statistics).
//
//   DYCORS  (candidate_dycors.py:73-86): coordinate j of `subset` is perturbed with probability prob_perturb; a row with
//           no perturbed coordinate gets one at randint(0, |subset| - 1) -- upper bound exclusive, so never the last one
//           (reference behaviour, kept); |subset| == 1 perturbs always.
//   SRBF    (candidate_srbf.py:108-113): every coordinate of `subset` is perturbed.
//   uniform (candidate_uniform.py:44-45): coordinates of `subset` are uniform in [lb, ub].
//   perturbation = truncated normal around xbest: loc xbest, scale sigma, support [lb, ub], drawn by inverse CDF;
//   integer variables are rounded and clipped (utils.py:68-93).
//
// RNG: Philox4x32-10, key = seed, counter = (row, draw index): stateless, reproducible for a given seed regardless of
// the launch geometry.
#include "common.cuh"

namespace b200rbf {
namespace {

struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t ka, uint32_t kb) const {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ ka, n1 = lo1, n2 = hi0 ^ c[3] ^ kb, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // two uniform doubles in (0, 1) for (row, draw)
  __device__ __forceinline__ void draw(unsigned long long row, uint32_t idx, double& u0, double& u1) const {
    uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), idx, 0x9E3779B9u};
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, ka, kb);
      ka += 0x9E3779B9u;
      kb += 0xBB67AE85u;
    }
    const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1], b = ((unsigned long long)c[2] << 32) | c[3];
    u0 = ((double)(a >> 11) + 0.5) * 1.1102230246251565e-16;  // 2^-53
    u1 = ((double)(b >> 11) + 0.5) * 1.1102230246251565e-16;
  }
};

struct GenArgs {
  int kind;  // 0 DYCORS, 1 SRBF, 2 uniform
  long long m;
  int d, nsub;
  const double* xbest;  // d
  const double* lb;     // d
  const double* ub;     // d
  const double* sigma;  // d
  const int* subset;    // nsub coordinate indices
  const unsigned char* int_mask;  // d, or nullptr
  double prob_perturb;
  unsigned long long seed;
  double* cand;  // m x d
};

// One WARP per row, lane = coordinate of `subset` (round 2; the first version gave a row to one thread, which walked its
// coordinates one after the other: a strategy's 100 d rows are two dozen CTAs and d sequential inverse-CDF
// evaluations, 45 us at d = 30 -- the largest kernel of a proposal).  The CDF bounds of the truncation depend on the
// coordinate only and are computed once per CTA.  Counter-based RNG: the values are those of the first version, bit
// for bit.
constexpr int kGenWarps = 8;        // rows in flight per CTA
constexpr int kGenRowsPerWarp = 2;  // rows a warp handles one after the other
__global__ void __launch_bounds__(32 * kGenWarps) generate_kernel(GenArgs a) {
  extern __shared__ double gs[];
  double* s_ca = gs;           // nsub: Phi((lb - xbest) / sigma)
  double* s_cb = gs + a.nsub;  // nsub: Phi((ub - xbest) / sigma)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (a.kind != 2) {
    for (int j = threadIdx.x; j < a.nsub; j += blockDim.x) {
      const int k = a.subset[j];
      const double xb = a.xbest[k], s = a.sigma[k];
      s_ca[j] = normcdf((a.lb[k] - xb) / s);
      s_cb[j] = normcdf((a.ub[k] - xb) / s);
    }
  }
  __syncthreads();
  Philox rng{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  for (int rr = 0; rr < kGenRowsPerWarp; ++rr) {
    const long long row = ((long long)blockIdx.x * kGenWarps + warp) * kGenRowsPerWarp + rr;
    if (row >= a.m) break;  // warp-uniform
    double* out = a.cand + row * a.d;
    for (int k = lane; k < a.d; k += 32) out[k] = a.xbest[k];  // np.multiply(np.ones((num_cand, dim)), xbest)
    __syncwarp();  // a perturbed coordinate may be written by another lane than the one that filled it
    int forced = -1;
    if (a.kind == 0 && a.nsub > 1) {
      // does any coordinate get perturbed?  (first uniform of every draw is the mask decision)
      bool any = false;
      for (int j0 = 0; j0 < a.nsub; j0 += 32) {
        const int j = j0 + lane;
        bool mine = false;
        if (j < a.nsub) {
          double u0, u1;
          rng.draw((unsigned long long)row, (uint32_t)j, u0, u1);
          mine = u0 < a.prob_perturb;
        }
        any |= __any_sync(0xffffffffu, mine);
      }
      if (!any) {
        double u0, u1;
        rng.draw((unsigned long long)row, 0x80000000u, u0, u1);
        forced = min((int)(u0 * (a.nsub - 1)), a.nsub - 2);  // np.random.randint(0, len(subset) - 1): last one never chosen
      }
    }
    for (int j = lane; j < a.nsub; j += 32) {
      const int k = a.subset[j];
      double u0, u1;
      rng.draw((unsigned long long)row, (uint32_t)j, u0, u1);
      const double lo = a.lb[k], hi = a.ub[k];
      double x;
      if (a.kind == 2) {
        x = lo + (hi - lo) * u1;  // np.random.uniform(lb, ub)
      } else {
        const bool perturb = a.kind == 1 || a.nsub == 1 || u0 < a.prob_perturb || j == forced;
        if (!perturb) continue;
        const double xb = a.xbest[k], s = a.sigma[k];
        // truncnorm(a=(lo-xb)/s, b=(hi-xb)/s, loc=xb, scale=s) by inverse CDF; lo <= xb <= hi so 0 is inside [a, b]
        const double ca = s_ca[j], cb = s_cb[j];
        x = xb + s * normcdfinv(ca + u1 * (cb - ca));
        x = fmin(fmax(x, lo), hi);
      }
      if (a.int_mask != nullptr && a.int_mask[k]) x = fmin(fmax(rint(x), lo), hi);  // np.round (half to even) + clip
      out[k] = x;
    }
    if (a.int_mask != nullptr) {  // round_vars also rounds untouched integer coordinates (xbest is integral already)
      __syncwarp();
      for (int k = lane; k < a.d; k += 32)
        if (a.int_mask[k]) out[k] = fmin(fmax(rint(out[k]), a.lb[k]), a.ub[k]);
    }
  }
}

}  // namespace

// small host arrays -> device (through the handle's tmp_box buffer), then one launch
int run_generate(b200rbf_ctx* h, int kind, long long m, int d, const double* xbest, const double* lb, const double* ub,
                 const double* sigma, double prob_perturb, const int* subset, int nsub, const unsigned char* int_mask,
                 unsigned long long seed, double* cand_dev) {
  const size_t nd = (size_t)d;
  const size_t bytes = 4 * nd * sizeof(double) + (size_t)nsub * sizeof(int) + nd + 64;
  B2_TRY(h->gen_buf.reserve(bytes));
  B2_TRY(h->gen_pin.reserve(bytes));
  B2_CUDA(cudaStreamSynchronize(h->stream));  // the pinned scratch may be in flight from the previous call
  char* hp = reinterpret_cast<char*>(h->gen_pin.p);
  double* hd = reinterpret_cast<double*>(hp);
  memcpy(hd, xbest, nd * 8);
  memcpy(hd + nd, lb, nd * 8);
  memcpy(hd + 2 * nd, ub, nd * 8);
  memcpy(hd + 3 * nd, sigma, nd * 8);
  int* hs = reinterpret_cast<int*>(hd + 4 * nd);
  for (int j = 0; j < nsub; ++j) hs[j] = subset ? subset[j] : j;
  unsigned char* hm = reinterpret_cast<unsigned char*>(hs + nsub);
  bool any_int = false;
  for (int k = 0; k < d; ++k) {
    hm[k] = int_mask ? int_mask[k] : 0;
    any_int |= hm[k] != 0;
  }
  B2_CUDA(cudaMemcpyAsync(h->gen_buf.p, hp, bytes, cudaMemcpyHostToDevice, h->stream));
  char* dp = reinterpret_cast<char*>(h->gen_buf.p);
  GenArgs a;
  a.kind = kind; a.m = m; a.d = d; a.nsub = nsub;
  a.xbest = reinterpret_cast<double*>(dp);
  a.lb = a.xbest + nd; a.ub = a.xbest + 2 * nd; a.sigma = a.xbest + 3 * nd;
  a.subset = reinterpret_cast<int*>(dp + 4 * nd * sizeof(double));
  a.int_mask = any_int ? reinterpret_cast<unsigned char*>(dp + 4 * nd * sizeof(double) + (size_t)nsub * sizeof(int)) : nullptr;
  a.prob_perturb = prob_perturb; a.seed = seed; a.cand = cand_dev;
  const long long rows_per_cta = kGenWarps * kGenRowsPerWarp;
  generate_kernel<<<(unsigned)((m + rows_per_cta - 1) / rows_per_cta), 32 * kGenWarps, 2 * (size_t)nsub * sizeof(double), h->stream>>>(a);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200rbf
