// Accuracy of the table-driven half-logarithm used by the DMMA score kernel's thin-plate-spline epilogue
// (score_mma.cuh: tps_half_log) against 0.5 * log(): max absolute and relative error over random doubles.
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include "../pysot_b200/csrc/score_mma.cuh"
namespace b200rbf { void set_error(const char*, ...) {} }
using namespace b200rbf;

__global__ void probe(const double2* __restrict__ table, unsigned long long seed, int lo_exp, int hi_exp, double* out) {
  __shared__ double2 tab[kTpsLogEntries];
  for (int i = threadIdx.x; i < kTpsLogEntries; i += blockDim.x) tab[i] = table[i];
  __syncthreads();
  curandStatePhilox4_32_10_t st;
  curand_init(seed, blockIdx.x * blockDim.x + threadIdx.x, 0, &st);
  double ea = 0, er = 0;
  for (int it = 0; it < 4096; ++it) {
    const double m = 1.0 + curand_uniform_double(&st);
    const int ex = lo_exp + (int)(curand(&st) % (unsigned)(hi_exp - lo_exp + 1));
    const double x = ldexp(m, ex);
    const double got = tps_half_log(x, tab), want = 0.5 * log(x);
    ea = fmax(ea, fabs(got - want));
    er = fmax(er, fabs(got - want) / fmax(fabs(want), 1e-300));
  }
  for (int o = 16; o; o >>= 1) { ea = fmax(ea, __shfl_xor_sync(~0u, ea, o)); er = fmax(er, __shfl_xor_sync(~0u, er, o)); }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(ea));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(er));
  }
}

int main() {
  std::vector<double> h(2 * kTpsLogEntries);
  tps_log_table_host(h.data());
  double2* tab; cudaMalloc(&tab, h.size() * 8); cudaMemcpy(tab, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  double* d; cudaMalloc(&d, 16);
  const int ranges[4][2] = {{-1, 0}, {-8, 8}, {-104, -90}, {-60, 12}};
  for (auto& r : ranges) {
    cudaMemset(d, 0, 16);
    probe<<<592, 256>>>(tab, 77, r[0], r[1], d);
    double o[2]; cudaMemcpy(o, d, 16, cudaMemcpyDeviceToHost);
    printf("x in 2^[%d,%d]: max |err| %.3e   max rel err %.3e  (%s)\n", r[0], r[1] + 1, o[0], o[1], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
