// Tiling sweep for the DMMA score kernel (development tool).  Built with different -DB2_MMA_MB / -DB2_MMA_MINB by
// tools/build_tune_mma.sh; times score_mma_kernel<K4, cubic> on BASELINE config #4's shape and prints pairs / s.
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#ifndef TUNE_KERN
#define TUNE_KERN 0
#endif
#ifndef TUNE_AUG
#define TUNE_AUG false
#endif
#include "../pysot_b200/csrc/score_mma.cuh"
namespace b200rbf {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
}
using namespace b200rbf;
#ifndef TUNE_D
#define TUNE_D 50
#endif
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main(int argc, char** argv) {
  const long long m = argc > 1 ? atoll(argv[1]) : (1 << 18);
  const int n = argc > 2 ? atoi(argv[2]) : 20000;
  const int reps = argc > 3 ? atoi(argv[3]) : 5;
  const int d = TUNE_D, ds = mma_stride(d), npad = pad_rows(n);
  constexpr int K4 = (TUNE_D + 3) / 4;
  std::vector<double> hc((size_t)m * d), hp((size_t)npad * ds, 0.0), hw(npad), hn(npad), ht(d + 1, 0.1), hb(2 * pad_dim(d));
  srand(1);
  for (auto& v : hc) v = rand() / (double)RAND_MAX;
  for (int i = 0; i < npad; ++i) { double s = 0; for (int k = 0; k < d; ++k) { double v = rand() / (double)RAND_MAX; hp[(size_t)i * ds + k] = v; s += v * v; } hn[i] = s; if (TUNE_AUG) { hp[(size_t)i * ds + d] = 1.0; hp[(size_t)i * ds + d + 1] = s; } hw[i] = rand() / (double)RAND_MAX - 0.5; }
  for (int k = 0; k < pad_dim(d); ++k) { hb[k] = 0.0; hb[pad_dim(d) + k] = 1.0; }
  double *cand, *pts, *coef, *nrm, *tailc, *box, *fv, *dm, *parts;
  const int nparts = (int)((m + kMmaRowsPerCta - 1) / kMmaRowsPerCta);
  CK(cudaMalloc(&cand, hc.size() * 8)); CK(cudaMalloc(&pts, hp.size() * 8)); CK(cudaMalloc(&coef, hw.size() * 8));
  CK(cudaMalloc(&nrm, hn.size() * 8)); CK(cudaMalloc(&tailc, ht.size() * 8)); CK(cudaMalloc(&box, hb.size() * 8));
  CK(cudaMalloc(&fv, m * 8)); CK(cudaMalloc(&dm, m * 8)); CK(cudaMalloc(&parts, (size_t)4 * nparts * 8));
  CK(cudaMemcpy(cand, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(pts, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(coef, hw.data(), hw.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(nrm, hn.data(), hn.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(tailc, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(box, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
  b200rbf_ctx ctx;
  CK(cudaStreamCreate(&ctx.stream));
  MmaArgs ma; memset(&ma, 0, sizeof(ma));
  ScoreArgs& a = ma.s;
  a.cand = cand; a.m = m; a.d = d; a.lb = box; a.range = box + pad_dim(d); a.pts = pts; a.coef = coef; a.tailc = tailc;
  a.npad = npad; a.tail = 0; a.dscale = 1.0; a.min_mode = 1; a.fvals = fv; a.dmerit = dm; a.parts = parts; a.nparts = nparts;
  ma.cnorm = nrm; ma.ds = ds; ma.aug = TUNE_AUG; ma.split = 1;
  { std::vector<double> ht2(2 * kTpsLogEntries); tps_log_table_host(ht2.data()); double* tt; CK(cudaMalloc(&tt, ht2.size() * 8)); CK(cudaMemcpy(tt, ht2.data(), ht2.size() * 8, cudaMemcpyHostToDevice)); ma.tps_table = reinterpret_cast<const double2*>(tt); }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&]() { return launch_score_mma_inst<K4, TUNE_KERN, TUNE_AUG>(&ctx, ma); };
  for (int i = 0; i < 2; ++i) if (run() != 0) return 1;
  CK(cudaStreamSynchronize(ctx.stream));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0, ctx.stream)); if (run() != 0) return 1; CK(cudaEventRecord(e1, ctx.stream)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
  }
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, score_mma_kernel<K4, TUNE_KERN, TUNE_AUG>, kMmaThreads, score_mma_smem_bytes(ds, TUNE_KERN == 1));
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, score_mma_kernel<K4, TUNE_KERN, TUNE_AUG>);
  const double pairs = (double)m * n;
  printf("mma d=%d mb=%d minb=%d jb=%d regs=%d ctas/sm=%d  best %.3f ms  %.4e pairs/s  %.2f TFLOP/s at %d flops/pair\n", d, kMmaMB,
         kMmaMinBlocks, kMmaJB, fa.numRegs, occ, best, pairs / (best * 1e-3), pairs * (3.0 * d + 6) / (best * 1e-3) / 1e12, 3 * d + 6);
  return 0;
}
