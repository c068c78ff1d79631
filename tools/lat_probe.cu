// Dependent-issue latencies seen by a single warp on sm_100a: DFMA, LDS.64 -> use, MUFU.RCP64H, __syncthreads with 16 warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  __shared__ double sm[1024];
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  double x = x0 + tid;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { x = fma(x, 1.0000001, 0.5); x = fma(x, 0.9999999, -0.5); x = fma(x, 1.0000001, 0.5); x = fma(x, 0.9999999, -0.5); }
  long long t1 = clock64();
  int idx = tid;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { idx = (int)sm[idx & 1023]; idx = (int)sm[idx & 1023]; idx = (int)sm[idx & 1023]; idx = (int)sm[idx & 1023]; }
  long long t2 = clock64();
  double y = x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y)); }
  long long t3 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  long long t4 = clock64();
  if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
  out[tid] = x + idx + y;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
  for (int threads : {32, 512}) {
    k<<<1, threads>>>(o, c, 1.0); k<<<1, threads>>>(o, c, 1.0);
    long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("%3d threads: DFMA dependent %.1f cycles, LDS.64 + F2I dependent %.1f, rcp.approx.f64 dependent %.1f, __syncthreads %.1f  (%s)\n",
           threads, h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
