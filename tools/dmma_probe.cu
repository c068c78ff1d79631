// FP64 pipe probes (development tool): DMMA m8n8k4 and DFMA throughput as a function of warps per SM and
// independent chains per warp, to size the GEMM / score kernels' occupancy.  Prints TFLOP/s per configuration.
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma(double* out, int iters) {
  double acc[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i][0] = acc[i][1] = 1.0;
  const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-15;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void k_dfma(double* out, int iters) {
  double f[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) f[i] = 1.0 + i + threadIdx.x * 1e-9;
  const double m = 1.0 + 1e-12, c = 1e-15;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) f[i] = fma(f[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

template <class K>
static double time_kernel(K kern, int blocks, int threads, int iters, double* sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(sink, iters / 10 + 1);
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(sink, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms * 1e-3;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink;
  cudaMalloc(&sink, 64);
  const int iters = 20000;
  printf("SMs %d\n", sms);
  const int wps[] = {4, 8, 12, 16, 32, 64};
  for (int w : wps) {
    const int threads = 32 * (w < 32 ? w : 32), blocks = sms * (w < 32 ? 1 : w / 32);
    const double warps = (double)blocks * threads / 32;
    double t;
#define RUN_DMMA(CH)                                                                                   \
  t = time_kernel(k_dmma<CH>, blocks, threads, iters, sink);                                           \
  printf("dmma warps/SM %2d chains %2d : %7.2f TFLOP/s\n", w, CH, warps * CH * 512.0 * iters / t / 1e12);
    RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8) RUN_DMMA(16) RUN_DMMA(32)
#define RUN_DFMA(CH)                                                                                   \
  t = time_kernel(k_dfma<CH>, blocks, threads, iters, sink);                                           \
  printf("dfma warps/SM %2d chains %2d : %7.2f TFLOP/s\n", w, CH, warps * 32 * CH * 2.0 * iters / t / 1e12);
    RUN_DFMA(1) RUN_DFMA(2) RUN_DFMA(4) RUN_DFMA(8) RUN_DFMA(16)
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
