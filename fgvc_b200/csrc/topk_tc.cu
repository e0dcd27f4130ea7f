// K1 (tcgen05 engine) -- placeholder until the tensor-core kernel lands.
#include "common.cuh"
namespace fgvc {
bool tc_supported(int, int, int, int) { return false; }
int launch_affinity_topk_tc(const float*, int, int, int, const fgvc_job*, int, const int32_t*, int, int, int, int,
                            float*, int32_t*, cudaStream_t) {
  set_error("tcgen05 engine not built");
  return FGVC_ERR_UNSUPPORTED;
}
}  // namespace fgvc
