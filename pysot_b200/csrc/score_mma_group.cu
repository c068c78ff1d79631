// Instantiations of the DMMA score kernel; compiled 4 times with -DB2_MGROUP=0..3.  Group g covers K4 = 4g+1 .. 4g+4
// (K4 = ceil(d / 4) k-steps of the m8n8k4 MMA).
#include "score_mma.cuh"

#ifndef B2_MGROUP
#error "compile with -DB2_MGROUP=<0..3>"
#endif
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)

namespace b200rbf {

int B2_CAT(launch_score_mma_group, B2_MGROUP)(b200rbf_ctx* h, const MmaArgs& ma, int k4, int kernel) {
  switch (k4 - 4 * B2_MGROUP) {
    case 1: return launch_score_mma_k4<4 * B2_MGROUP + 1>(h, ma, kernel);
    case 2: return launch_score_mma_k4<4 * B2_MGROUP + 2>(h, ma, kernel);
    case 3: return launch_score_mma_k4<4 * B2_MGROUP + 3>(h, ma, kernel);
    case 4: return launch_score_mma_k4<4 * B2_MGROUP + 4>(h, ma, kernel);
  }
  set_error("internal: k4 %d not in MMA group %d", k4, B2_MGROUP);
  return B200RBF_EINVAL;
}

}  // namespace b200rbf
