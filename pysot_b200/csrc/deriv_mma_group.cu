// Instantiations of the DMMA gradient kernel; compiled 4 times with -DB2_XGROUP=0..3 (K4 = 4g+1 .. 4g+4).
#include "deriv_mma.cuh"

#ifndef B2_XGROUP
#error "compile with -DB2_XGROUP=<0..3>"
#endif
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)

namespace b200rbf {

int B2_CAT(launch_deriv_mma_group, B2_XGROUP)(b200rbf_ctx* h, const DerivMmaArgs& a, int k4, int kernel, bool aug) {
  switch (k4 - 4 * B2_XGROUP) {
    case 1: return launch_deriv_mma_k4<4 * B2_XGROUP + 1>(h, a, kernel, aug);
    case 2: return launch_deriv_mma_k4<4 * B2_XGROUP + 2>(h, a, kernel, aug);
    case 3: return launch_deriv_mma_k4<4 * B2_XGROUP + 3>(h, a, kernel, aug);
    case 4: return launch_deriv_mma_k4<4 * B2_XGROUP + 4>(h, a, kernel, aug);
  }
  set_error("internal: k4 %d not in gradient MMA group %d", k4, B2_XGROUP);
  return B200RBF_EINVAL;
}

}  // namespace b200rbf
