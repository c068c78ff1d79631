// Accuracy of the Newton-corrected MUFU.RSQ64H square root used by the DMMA score kernel (score_mma.cuh) against the
// correctly rounded sqrt(): max relative error of the seed, of one correction and of two, over random doubles.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <curand_kernel.h>

__device__ __forceinline__ void variants(double x, double& seed, double& s1, double& s2) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = __longlong_as_double(__double_as_longlong(y) - (1LL << 52));
  const double g = x * y;
  seed = g;
  s1 = fma(fma(-g, g, x), h, g);
  s2 = fma(fma(-s1, s1, x), h, s1);
}

__global__ void probe(unsigned long long seed, int lo_exp, int hi_exp, double* out) {
  curandStatePhilox4_32_10_t st;
  curand_init(seed, blockIdx.x * blockDim.x + threadIdx.x, 0, &st);
  double e0 = 0, e1 = 0, e2 = 0;
  for (int it = 0; it < 4096; ++it) {
    const double m = 1.0 + curand_uniform_double(&st);
    const int ex = lo_exp + (int)(curand(&st) % (unsigned)(hi_exp - lo_exp + 1));
    const double x = ldexp(m, ex);
    double a, b, c;
    variants(x, a, b, c);
    const double r = sqrt(x);
    e0 = fmax(e0, fabs(a - r) / r); e1 = fmax(e1, fabs(b - r) / r); e2 = fmax(e2, fabs(c - r) / r);
  }
  for (int o = 16; o; o >>= 1) {
    e0 = fmax(e0, __shfl_xor_sync(~0u, e0, o)); e1 = fmax(e1, __shfl_xor_sync(~0u, e1, o)); e2 = fmax(e2, __shfl_xor_sync(~0u, e2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e0));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e1));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(e2));
  }
}

int main() {
  double* d; cudaMalloc(&d, 24);
  const int ranges[3][2] = {{-40, 10}, {-1000, -900}, {900, 1000}};
  for (auto& r : ranges) {
    cudaMemset(d, 0, 24);
    probe<<<592, 256>>>(1234, r[0], r[1], d);
    double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("x in 2^[%d,%d]: max rel err  seed %.3e (2^%.1f)   1 correction %.3e   2 corrections %.3e  (%s)\n", r[0], r[1], h[0],
           log2(h[0]), h[1], h[2], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
