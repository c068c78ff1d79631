// K3: gradient of the RBF interpolant, RBFInterpolant.predict_deriv (rbf.py:187-215):
//   grad_k(x) = ( dP_k . lambda + sum_j (u_k - u_jk) * (phi'(r_j) * c_j) / r_j ) / (ub_k - lb_k),   r_j clamped to >= eps
// Query counts are small (the reference loops over them in Python), so the center axis is what gets
// parallelised: grid = (center slabs, queries).  Phase 1 (thread per center) computes the scalar
// s_j = phi'(r_j) c_j / r_j; phase 2 (thread per coordinate x center slice) accumulates (u_k - u_jk) s_j in
// the reference's operation order; a second tiny kernel adds the slab partials in slab order, the tail
// derivative (tails/linear_tail.py:46, constant_tail.py:44) and divides by the box width (rbf.py:215).
#include "deriv_mma.cuh"

namespace b200rbf {
int ensure_tps_table(b200rbf_ctx* h);  // api.cu
namespace {

constexpr int kDerivThreads = 256;
constexpr int kSlab = 512;
constexpr long long kDerivTileMinQueries = 4096;  // below this a thread per query cannot fill 148 SMs

__device__ __forceinline__ double dphi_times_c_over_r(int kern, double r, double c) {
  // kernels/cubic_kernel.py:37 3 r^2 ; tps_kernel.py:41-42 r (1 + 2 log r) ; linear_kernel.py:40 1
  // rbf.py:212: np.multiply(kernel.deriv(dss), c) / dss
  double dp;
  if (kern == B200RBF_KERNEL_CUBIC) {
    dp = __dmul_rn(3.0, __dmul_rn(r, r));
  } else if (kern == B200RBF_KERNEL_TPS) {
    dp = __dmul_rn(r, __dadd_rn(1.0, __dmul_rn(2.0, log(r))));
  } else {
    dp = 1.0;
  }
  return __ddiv_rn(__dmul_rn(dp, c), r);
}

template <bool EXACT>
__global__ void __launch_bounds__(kDerivThreads) deriv_partial_kernel(
    const double* __restrict__ x, int d, const double* __restrict__ lb, const double* __restrict__ range,
    const double* __restrict__ centers, int dpad, const double* __restrict__ coef, int n, int kern,
    double* __restrict__ partial /* [m][nslabs][d] */) {
  extern __shared__ double sm[];
  double* s_u = sm;              // d
  double* s_s = s_u + d;         // kSlab
  double* s_acc = s_s + kSlab;   // kDerivThreads
  const int q = blockIdx.y, slab = blockIdx.x, nslabs = gridDim.x, tid = threadIdx.x;
  const int j0 = slab * kSlab;
  const int cnt = min(kSlab, n - j0);
  for (int k = tid; k < d; k += blockDim.x)
    s_u[k] = __ddiv_rn(__dsub_rn(x[(long long)q * d + k], lb[k]), range[k]);  // utils.py:33
  __syncthreads();
  const double eps = 2.220446049250313e-16;
  for (int j = tid; j < cnt; j += blockDim.x) {
    const double* cp = centers + (size_t)(j0 + j) * dpad;
    double acc = 0.0;
    for (int k = 0; k < d; ++k) {
      const double t = __ldg(cp + k) - s_u[k];  // cdist(self._X, xx): center minus query (rbf.py:200)
      acc = EXACT ? __dadd_rn(acc, __dmul_rn(t, t)) : fma(t, t, acc);
    }
    double r = sqrt(acc);
    if (r < eps) r = eps;  // rbf.py:201
    s_s[j] = dphi_times_c_over_r(kern, r, coef[j0 + j]);
  }
  __syncthreads();
  // phase 2: thread = (coordinate k, slice sl); lanes run over k so the row-major centers are read coalesced
  const int kd = min(d, (int)blockDim.x);
  const int nsl = blockDim.x / kd;
  const int kk = tid % kd, sl = tid / kd;
  for (int kb = 0; kb < d; kb += kd) {
    const int k = kb + kk;
    double acc = 0.0;
    if (k < d && sl < nsl) {
      const double uk = s_u[k];
      for (int j = sl; j < cnt; j += nsl) {
        const double diff = __dsub_rn(uk, __ldg(centers + (size_t)(j0 + j) * dpad + k));  // rbf.py:209-210
        acc = __dadd_rn(acc, __dmul_rn(diff, s_s[j]));                                      // rbf.py:212-213
      }
    }
    s_acc[tid] = acc;
    __syncthreads();
    if (sl == 0 && k < d) {
      double tot = acc;
      for (int s2 = 1; s2 < nsl; ++s2) tot += s_acc[s2 * kd + kk];
      partial[((size_t)q * nslabs + slab) * d + k] = tot;
    }
    __syncthreads();
  }
}

__global__ void deriv_final_kernel(const double* __restrict__ partial, int nslabs, int d, long long m,
                                   const double* __restrict__ tailc, int tail, const double* __restrict__ range,
                                   double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m * d) return;
  const long long q = e / d;
  const int k = (int)(e - q * d);
  double g = (tail == B200RBF_TAIL_LINEAR) ? tailc[1 + k] : 0.0;  // dP . lambda (rbf.py:207-208)
  double s = 0.0;
  for (int sl = 0; sl < nslabs; ++sl) s += partial[((size_t)q * nslabs + sl) * d + k];
  out[e] = __ddiv_rn(g + s, range[k]);  // rbf.py:215
}

}  // namespace

// scratch must hold m * nslabs * d doubles, nslabs = ceil(n / kSlab)
int launch_deriv(b200rbf_ctx* h, const double* x, long long m, const double* lb, const double* range, double* out,
                 double* scratch) {
  const Model& M = h->model;
  if (m >= kDerivTileMinQueries && M.dpad <= kMaxRegDim && M.kernel != B200RBF_KERNEL_LINEAR && h->mma_dist) {
    // enough queries to fill the GPU: distances AND the scaled accumulations as two chained DMMA GEMMs (deriv_mma.cuh)
    DerivMmaArgs a;
    a.x = x; a.m = m; a.d = M.d; a.lb = lb; a.range = range;
    a.pts = M.aug ? M.centers_m.as<double>() : M.centers_x.as<double>();
    a.cnorm = M.cnorm.as<double>(); a.coef = M.coef.as<double>(); a.tailc = M.tailc.as<double>();
    a.ds = M.ds; a.npad = M.npad; a.tail = M.tail; a.out = out;
    a.tps_table = nullptr;
    if (M.kernel == B200RBF_KERNEL_TPS) {
      B2_TRY(ensure_tps_table(h));
      a.tps_table = h->tps_table.as<double2>();
    }
    const int k4 = (M.d + 3) / 4;
    switch ((k4 - 1) / 4) {
      case 0: return launch_deriv_mma_group0(h, a, k4, M.kernel, M.aug);
      case 1: return launch_deriv_mma_group1(h, a, k4, M.kernel, M.aug);
      case 2: return launch_deriv_mma_group2(h, a, k4, M.kernel, M.aug);
      default: return launch_deriv_mma_group3(h, a, k4, M.kernel, M.aug);
    }
  }
  if (m >= kDerivTileMinQueries && M.dpad <= kMaxRegDim) {
    // the linear kernel (s = w / r is unbounded at coincident points) and B200RBF_OPT_MMA_DIST = 0: one thread per
    // query, direct differences (deriv_tile.cuh)
    DerivTileArgs a;
    a.x = x; a.m = m; a.d = M.d; a.lb = lb; a.range = range;
    a.pts = M.centers_x.as<double>(); a.coef = M.coef.as<double>(); a.tailc = M.tailc.as<double>();
    a.ds = M.ds; a.npad = M.npad; a.tail = M.tail; a.out = out;
    a.tps_table = nullptr;
    if (M.kernel == B200RBF_KERNEL_TPS) {
      B2_TRY(ensure_tps_table(h));
      a.tps_table = h->tps_table.as<double2>();
    }
    const int k4 = (M.d + 3) / 4;
    switch ((k4 - 1) / 4) {
      case 0: return launch_deriv_tile_group0(h, a, k4, M.kernel);
      case 1: return launch_deriv_tile_group1(h, a, k4, M.kernel);
      case 2: return launch_deriv_tile_group2(h, a, k4, M.kernel);
      default: return launch_deriv_tile_group3(h, a, k4, M.kernel);
    }
  }
  const int nslabs = (M.n + kSlab - 1) / kSlab;
  const size_t smem = ((size_t)M.d + kSlab + kDerivThreads) * sizeof(double);
  auto kern = h->exact_dist ? deriv_partial_kernel<true> : deriv_partial_kernel<false>;
  if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long kMaxY = 65535;
  for (long long q0 = 0; q0 < m; q0 += kMaxY) {
    const long long mq = (m - q0 < kMaxY) ? (m - q0) : kMaxY;
    dim3 grid((unsigned)nslabs, (unsigned)mq);
    kern<<<grid, kDerivThreads, smem, h->stream>>>(x + q0 * M.d, M.d, lb, range, M.centers.as<double>(), M.dpad,
                                                   M.coef.as<double>(), M.n, M.kernel,
                                                   scratch + (size_t)q0 * nslabs * M.d);
    h->launches++;
    B2_CUDA(cudaGetLastError());
  }
  const long long total = m * M.d;
  deriv_final_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(scratch, nslabs, M.d, m,
                                                                             M.tailc.as<double>(), M.tail, range, out);
  h->launches++;
  B2_CUDA(cudaGetLastError());
  return 0;
}

int deriv_scratch_doubles(const b200rbf_ctx* h, long long m, size_t* out) {
  const Model& M = h->model;
  const int nslabs = (M.n + kSlab - 1) / kSlab;
  const bool tile = m >= kDerivTileMinQueries && M.dpad <= kMaxRegDim;  // the tile kernel needs no partials
  *out = tile ? 0 : (size_t)m * nslabs * M.d;
  return 0;
}

}  // namespace b200rbf
