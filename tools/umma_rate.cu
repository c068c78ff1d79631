// Stage-2 probe: issue rate of tcgen05.mma.kind::i8 M128 x N x K32 with both operands in shared memory, as a function of
// N and of the CTAs per SM -- decides the tile shape of the int8 scoring experiment.
//   build/tune/umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int N, int COLS, int NACC>
__global__ void __launch_bounds__(128) rate_kernel(int rounds, int per_round, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                // 6 slices x 128 rows x 128 B (two 64-B halves used)
  unsigned char* sB = smem + 6 * 128 * 128;  // 6 slices x N rows x 128 B
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 6 * 128 * 128 + 6 * N * 128; e += 128) smem[e] = (unsigned char)(e * 7 + 3);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;
  for (int r = 0; r < rounds; ++r) {
    if (tid == 0) {
      for (int q = 0; q < per_round; ++q) {
        const int w = q % NACC, i = q % 6, half = q & 1;
        const uint64_t da = make_desc(smem_u32(sA) + i * 128 * 128) + (uint64_t)(half * 2);
        const uint64_t db = make_desc(smem_u32(sB) + ((q + 3) % 6) * N * 128) + (uint64_t)(half * 2);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem + (uint32_t)(w * N)),
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS));
}
template <int N, int COLS, int NACC>
int run(int ctas_per_sm) {
  int* st;
  CK(cudaMalloc(&st, 4));
  CK(cudaMemset(st, 0, 4));
  const size_t smem = 6 * 128 * 128 + 6 * N * 128 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<N, COLS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rounds = 400, per = 42;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<N, COLS, NACC><<<148 * ctas_per_sm, 128, smem>>>(10, per, st);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  rate_kernel<N, COLS, NACC><<<148 * ctas_per_sm, 128, smem>>>(rounds, per, st);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int s = 0; cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
  const double mmas = (double)rounds * per * 148 * ctas_per_sm;
  const double ops = mmas * 2.0 * 128 * N * 32;
  printf("N=%3d cols=%3d nacc=%d CTAs/SM=%d: %.3f ms, %.1f ns per MMA per SM-slot, %.2f POPS, pairs/s bound for 42 MMAs per 128xN tile: %.3e (status %d)\n",
         N, COLS, NACC, ctas_per_sm, ms, ms * 1e6 / (rounds * per), ops / ms * 1e-12, mmas / 42 * 128 * N / (ms * 1e-3), s);
  cudaFree(st);
  return 0;
}
int main() {
  run<64, 512, 6>(1); run<64, 64, 1>(1); run<64, 128, 2>(1); run<128, 512, 4>(1); run<128, 128, 1>(1); run<256, 512, 2>(1); run<256, 256, 1>(1); run<256, 256, 1>(2); run<32, 32, 1>(1);
  return 0;
}
