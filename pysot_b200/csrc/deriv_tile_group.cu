// Instantiations of the tile-streaming gradient kernel; compiled 4 times with -DB2_DGROUP=0..3 (K4 = 4g+1 .. 4g+4).
#include "deriv_tile.cuh"

#ifndef B2_DGROUP
#error "compile with -DB2_DGROUP=<0..3>"
#endif
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)

namespace b200rbf {

int B2_CAT(launch_deriv_tile_group, B2_DGROUP)(b200rbf_ctx* h, const DerivTileArgs& a, int k4, int kernel) {
  switch (k4 - 4 * B2_DGROUP) {
    case 1: return launch_deriv_tile_k4<4 * B2_DGROUP + 1>(h, a, kernel);
    case 2: return launch_deriv_tile_k4<4 * B2_DGROUP + 2>(h, a, kernel);
    case 3: return launch_deriv_tile_k4<4 * B2_DGROUP + 3>(h, a, kernel);
    case 4: return launch_deriv_tile_k4<4 * B2_DGROUP + 4>(h, a, kernel);
  }
  set_error("internal: k4 %d not in gradient group %d", k4, B2_DGROUP);
  return B200RBF_EINVAL;
}

}  // namespace b200rbf
