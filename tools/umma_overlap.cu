// Does tcgen05.mma.kind::i8 run concurrently with FP64 work of other warps of the same CTA?  One warp issues 52-MMA rounds
// (M128 N64 K32, commit + wait per round), the other warps run: nothing / a DFMA loop / a DFMA loop with shared-memory
// loads / an integer loop, for the same number of rounds.  Prints the time of each mix.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
constexpr int N = 64, PER = 52, NACC = 7;
__global__ void __launch_bounds__(544) k(int rounds, int mode, int do_mma, int work, double* sink, int* status) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sA = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = sA + 2 * 128 * 128;
  double* sD = reinterpret_cast<double*>(sB + 2 * N * 128);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 2 * 128 * 128 + 2 * N * 128; e += 544) sA[e] = (unsigned char)(e * 7 + 3);
  for (int e = tid; e < 256; e += 544) sD[e] = 1.0 + e * 1e-9;
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
  if (warp == 0) {
    if (tid == 0 && do_mma) {
      uint32_t phase = 0;
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const uint64_t da = dA + (uint64_t)((q & 1) * 1024 + 2 * ((q >> 1) & 1));
          const uint64_t db = dB + (uint64_t)(((q >> 2) & 1) * (N * 8) + 2 * ((q >> 1) & 1));
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem + (uint32_t)((q % NACC) * N)),
                       "l"(da), "l"(db), "r"(idesc), "r"(q >= NACC ? 1u : 0u), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 24) && !ok; ++spin)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
        if (!ok) *status = 1;
        phase ^= 1;
      }
    }
  } else if (mode != 0) {
    // the other 16 warps: `work` iterations per round of the chosen instruction mix
    double a0 = 1.0 + tid * 1e-6, a1 = 1.5, a2 = 0.7, a3 = 1.1;
    long long i0 = tid, i1 = 3;
    for (int r = 0; r < rounds; ++r) {
      for (int w = 0; w < work; ++w) {
        if (mode == 1) {
          a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 0.9999999, 1e-9); a2 = fma(a2, 1.0000002, 1e-9); a3 = fma(a3, 0.9999998, 1e-9);
        } else if (mode == 2) {
          const double t = sD[(w + tid) & 255];
          a0 = fma(a0, t, 1e-9); a1 = fma(a1, t, 1e-9); a2 = fma(a2, t, 1e-9); a3 = fma(a3, t, 1e-9);
        } else {
          i0 = i0 * 3 + i1; i1 = i1 * 5 + i0; i0 ^= i1 >> 3; i1 += i0 << 1;
        }
      }
    }
    if (a0 + a1 + a2 + a3 + (double)(i0 + i1) == 123.456) sink[tid] = a0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
  int* st; double* sink;
  CK(cudaMalloc(&st, 4)); CK(cudaMemset(st, 0, 4)); CK(cudaMalloc(&sink, 8 * 1024));
  const size_t smem = 2 * 128 * 128 + 2 * N * 128 + 2048 + 2048;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rounds = 300;
  const char* names[] = {"idle", "DFMA loop", "DFMA + LDS loop", "integer loop"};
  for (int mode = 0; mode < 4; ++mode)
    for (int do_mma = 0; do_mma < 2; ++do_mma) {
      if (mode == 0 && !do_mma) continue;
      const int work = mode == 3 ? 400 : 200;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<148, 544, smem>>>(10, mode, do_mma, work, sink, st);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      k<<<148, 544, smem>>>(rounds, mode, do_mma, work, sink, st);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("other warps: %-16s MMAs %s: %.3f ms (%.2f us per round)\n", names[mode], do_mma ? "on " : "off", ms, ms * 1e3 / rounds);
    }
  int s = 0; cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
  printf("status %d\n", s);
  return 0;
}
