// Stage-1 probe for the Ozaki-style int8 scoring experiment ("int8 slices on tcgen05"):
// one CTA, D[128 x N] (int32, TMEM) = A[128 x K] (int8, smem, K-major, 128B swizzle) * B[N x K]^T with
// tcgen05.mma.kind::i8, read back with tcgen05.ld and compared with exact integer dot products on the host.
// It validates, in one go, the shared-memory descriptor, the manual 128-byte swizzle, the instruction descriptor,
// the commit -> mbarrier hand-off and the TMEM lane / column mapping the full kernel relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tune/umma_probe tools/umma_probe.cu && build/tune/umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);   \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kM = 128, kN = 64, kRowBytes = 128;  // one swizzle atom per row: up to 128 int8 of K
constexpr uint32_t kTmemCols = 64;

// byte offset of (row r, byte k) in a K-major tile whose rows are 128 B wide, 128B-swizzled (Swizzle<3,4,3>):
// 8-row groups of 1024 B; inside a group the 16-byte chunk index is XOR-ed with the row index
__host__ __device__ inline int sw128(int r, int k) { return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 4) ^ (r & 7)) & 7) << 4) + (k & 15); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major): 1
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8-row groups are 1024 B apart
  d |= (uint64_t)1 << 46;                         // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int K,
                                                    int32_t* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                       // 128 rows x 128 B = 16 KB
  unsigned char* sB = smem + kM * kRowBytes;      // 64 rows x 128 B = 8 KB
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < kM * kRowBytes; e += 128) sA[e] = 0;
  for (int e = tid; e < kN * kRowBytes; e += 128) sB[e] = 0;
  __syncthreads();
  for (int e = tid; e < kM * K; e += 128) sA[sw128(e / K, e % K)] = (unsigned char)A[e];
  for (int e = tid; e < kN * K; e += 128) sB[sw128(e / K, e % K)] = (unsigned char)B[e];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // generic-proxy writes to shared memory -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    // instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), K-major both, N >> 3, M >> 4
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
    for (int k = 0; k < K / 32; ++k) {
      const uint64_t dak = da + (uint64_t)((k * 32) >> 4), dbk = db + (uint64_t)((k * 32) >> 4);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
          "}\n" ::"r"(tmem),
          "l"(dak), "l"(dbk), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs (bounded: a wrong descriptor must not hang the box)
  {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(ok)
          : "r"(smem_u32(&mbar)), "r"(0)
          : "memory");
    }
    if (!ok && tid == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows of D), 64 columns, 16 at a time
  for (int c0 = 0; c0 < kN; c0 += 16) {
    uint32_t v[16];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * kN + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

int main() {
  int fails = 0;
  for (int K : {32, 64, 128}) {
    std::vector<int8_t> A(kM * K), B(kN * K);
    srand(K);
    for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
    int8_t *dA, *dB;
    int32_t* dD;
    int* dS;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, kM * kN * 4));
    CK(cudaMalloc(&dS, 4));
    CK(cudaMemset(dD, 0xff, kM * kN * 4));
    CK(cudaMemset(dS, 0, 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    const size_t smem = (kM + kN) * kRowBytes + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, K, dD, dS);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D(kM * kN);
    int st = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    int bad = 0, first = -1;
    for (int i = 0; i < kM; ++i)
      for (int j = 0; j < kN; ++j) {
        int ref = 0;
        for (int k = 0; k < K; ++k) ref += (int)A[i * K + k] * (int)B[j * K + k];
        if (ref != D[i * kN + j]) {
          if (first < 0) first = i * kN + j;
          ++bad;
        }
      }
    printf("K=%3d: status %d, %d / %d wrong", K, st, bad, kM * kN);
    if (bad) {
      int i = first / kN, j = first % kN, ref = 0;
      for (int k = 0; k < K; ++k) ref += (int)A[i * K + k] * (int)B[j * K + k];
      printf(" (first at row %d col %d: got %d, want %d; D[0][0..3] = %d %d %d %d)", i, j, D[first], ref, D[0], D[1], D[2], D[3]);
    }
    printf("\n");
    fails += bad != 0 || st != 0;
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  }
  printf(fails ? "UMMA PROBE FAILED\n" : "UMMA PROBE OK\n");
  return fails ? 2 : 0;
}
