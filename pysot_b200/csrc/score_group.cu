// One translation unit per group of padded dimensions; compiled 8 times with -DB2_GROUP=0..7 so the
// fully unrolled instantiations build in parallel.  Group g covers DPAD = 8g+2, 8g+4, 8g+6, 8g+8.
#include "score_kernels.cuh"

#ifndef B2_GROUP
#error "compile with -DB2_GROUP=<0..7>"
#endif
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)

namespace b200rbf {

int B2_CAT(launch_score_group, B2_GROUP)(b200rbf_ctx* h, const ScoreArgs& a, int dpad, int kernel, bool exact,
                                         bool with_phi) {
  switch (dpad - 8 * B2_GROUP) {
    case 2: return launch_score_dpad<8 * B2_GROUP + 2>(h, a, kernel, exact, with_phi);
    case 4: return launch_score_dpad<8 * B2_GROUP + 4>(h, a, kernel, exact, with_phi);
    case 6: return launch_score_dpad<8 * B2_GROUP + 6>(h, a, kernel, exact, with_phi);
    case 8: return launch_score_dpad<8 * B2_GROUP + 8>(h, a, kernel, exact, with_phi);
  }
  set_error("internal: dpad %d not in group %d", dpad, B2_GROUP);
  return B200RBF_EINVAL;
}

}  // namespace b200rbf
