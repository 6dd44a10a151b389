// tcgen05 / TMEM / TMA similarity kernel — placeholder until the tensor-core path lands.
#include "common.cuh"

namespace sslam {

size_t match_tc_extra_workspace(int, int, int, int, int) { return 0; }

int match_top2_tc(const void*, const void*, const int32_t*, int dtype, int, int, int, int, int32_t*,
                  float*, float*, u64*, void*, size_t, cudaStream_t) {
  set_error("match: dtype %d (tensor-core path) not built yet", dtype);
  return SSLAM_EUNSUPPORTED;
}

}  // namespace sslam
