// tcgen05.mma.kind::i8 M128 x N x K32 issue rate with a fully unrolled, address-arithmetic-free issue loop.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int N, int NACC, int PER>
__global__ void __launch_bounds__(128) rate_kernel(int rounds, int* status) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sA = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = sA + 2 * 128 * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 2 * 128 * 128 + 2 * N * 128; e += 128) sA[e] = (unsigned char)(e * 7 + 3);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t dA = make_desc(smem_u32(sA)), dB = make_desc(smem_u32(sB));
  uint32_t phase = 0;
  for (int r = 0; r < rounds; ++r) {
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const uint64_t da = dA + (uint64_t)((q & 1) * 1024 + 2 * ((q >> 1) & 1));
        const uint64_t db = dB + (uint64_t)(((q >> 2) & 1) * (N * 8) + 2 * ((q >> 1) & 1));
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem + (uint32_t)((q % NACC) * N)),
            "l"(da), "l"(db), "r"(idesc), "r"(q >= NACC ? 1u : 0u), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; ++spin)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
    if (!ok && tid == 0) *status = 1;
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N, int NACC, int PER>
int run() {
  int* st;
  CK(cudaMalloc(&st, 4));
  CK(cudaMemset(st, 0, 4));
  const size_t smem = 2 * 128 * 128 + 2 * N * 128 + 2048;
  CK(cudaFuncSetAttribute(rate_kernel<N, NACC, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rounds = 400;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<N, NACC, PER><<<148, 128, smem>>>(10, st);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  rate_kernel<N, NACC, PER><<<148, 128, smem>>>(rounds, st);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int s = 0; cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
  const double mmas = (double)rounds * PER * 148;
  printf("N=%3d nacc=%d %d MMAs per commit: %.1f ns per MMA, %.2f POPS (status %d)\n", N, NACC, PER, ms * 1e6 / (rounds * PER),
         mmas * 2.0 * 128 * N * 32 / ms * 1e-12, s);
  cudaFree(st);
  return 0;
}
int main() {
  run<64, 6, 42>(); run<64, 1, 42>(); run<128, 4, 42>(); run<256, 2, 42>(); run<256, 1, 42>(); run<64, 6, 168>(); run<32, 6, 42>(); run<16, 6, 42>();
  return 0;
}
