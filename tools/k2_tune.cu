// Tiling sweep for the fused score kernel (development tool, not part of libb200rbf.so).
// Built several times with different -DB2_SCORE_THREADS / -DB2_SCORE_MINB / -DB2_JT / -DB2_TILE_ROWS by
// tools/build_tune.sh; each binary times score_kernel<DPAD=50, cubic, fma> on synthetic data of BASELINE
// config #4's shape (d = 50, n = 20 000 centers) with CUDA events and prints candidate x center pairs / s.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "../pysot_b200/csrc/score_kernels.cuh"

namespace b200rbf {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
}  // namespace b200rbf
using namespace b200rbf;

#ifndef TUNE_DPAD
#define TUNE_DPAD 50
#endif
#ifndef TUNE_KERN
#define TUNE_KERN 0
#endif
#ifndef TUNE_EXACT
#define TUNE_EXACT false
#endif

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                     \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  const long long m = argc > 1 ? atoll(argv[1]) : (1 << 18);
  const int n = argc > 2 ? atoi(argv[2]) : 20000;
  const int reps = argc > 3 ? atoi(argv[3]) : 5;
  const int d = TUNE_DPAD, dpad = TUNE_DPAD, npad = pad_rows(n);
  std::vector<double> hc((size_t)m * d), hp((size_t)npad * dpad), hw(npad), ht(d + 1, 0.1), hb(2 * dpad);
  srand(1);
  for (auto& v : hc) v = rand() / (double)RAND_MAX;
  for (auto& v : hp) v = rand() / (double)RAND_MAX;
  for (auto& v : hw) v = rand() / (double)RAND_MAX - 0.5;
  for (int k = 0; k < dpad; ++k) {
    hb[k] = 0.0;
    hb[dpad + k] = 1.0;
  }
  double *cand, *pts, *coef, *tailc, *box, *fv, *dm, *parts;
  const int nparts = (int)((m + kScoreThreads - 1) / kScoreThreads);
  CK(cudaMalloc(&cand, hc.size() * 8));
  CK(cudaMalloc(&pts, hp.size() * 8));
  CK(cudaMalloc(&coef, hw.size() * 8));
  CK(cudaMalloc(&tailc, ht.size() * 8));
  CK(cudaMalloc(&box, hb.size() * 8));
  CK(cudaMalloc(&fv, m * 8));
  CK(cudaMalloc(&dm, m * 8));
  CK(cudaMalloc(&parts, (size_t)4 * nparts * 8));
  CK(cudaMemcpy(cand, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pts, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(coef, hw.data(), hw.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(tailc, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(box, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
  b200rbf_ctx ctx;
  CK(cudaStreamCreate(&ctx.stream));
  ScoreArgs a;
  memset(&a, 0, sizeof(a));
  a.cand = cand; a.m = m; a.d = d; a.lb = box; a.range = box + dpad; a.pts = pts; a.coef = coef; a.tailc = tailc;
  a.npad = npad; a.tail = 0; a.dscale = 1.0; a.min_mode = 1; a.fvals = fv; a.dmerit = dm; a.parts = parts;
  a.nparts = nparts; a.part_off = 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto run = [&]() { return launch_score_inst<TUNE_DPAD, TUNE_KERN, TUNE_EXACT, true>(&ctx, a); };
  for (int i = 0; i < 2; ++i)
    if (run() != 0) return 1;
  CK(cudaStreamSynchronize(ctx.stream));
  float best = 1e30f, tot = 0.f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0, ctx.stream));
    if (run() != 0) return 1;
    CK(cudaEventRecord(e1, ctx.stream));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
    tot += ms;
  }
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, score_kernel<TUNE_DPAD, TUNE_KERN, TUNE_EXACT, true>, kScoreThreads,
                                                score_smem_bytes<TUNE_DPAD>());
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, score_kernel<TUNE_DPAD, TUNE_KERN, TUNE_EXACT, true>);
  const double pairs = (double)m * n;
  printf("threads=%d minb=%d jt=%d tile=%d stages=%d kern=%d exact=%d regs=%d ctas/sm=%d  best %.3f ms  mean %.3f ms  "
         "%.4e pairs/s (best)  %.2f TFLOP/s at %d flops/pair\n",
         kScoreThreads, kScoreMinBlocks, kJT, kTileRows, kStages, TUNE_KERN, (int)TUNE_EXACT, fa.numRegs, occ, best,
         tot / reps, pairs / (best * 1e-3), pairs * (3.0 * d + 6) / (best * 1e-3) / 1e12, 3 * d + 6);
  return 0;
}
