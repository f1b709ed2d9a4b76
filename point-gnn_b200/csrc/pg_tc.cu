// tcgen05 (5th-gen tensor core) paths - placeholder until the UMMA kernels land.
#include "pg_common.cuh"

extern "C" int pg_tc_available(void) { return 0; }

namespace pg {

int fc_tc_bf16x3(const float*, int64_t, int, const float*, const float*, int, int, const float*, float*,
                 cudaStream_t) {
  set_error("tcgen05 fully-connected path not built yet");
  return PG_ERR_UNSUPPORTED;
}

int edge_mlp_max_tc(int, const float*, int, const float*, const float*, const int32_t*, const int32_t*,
                    const int32_t*, int64_t, int64_t, int64_t, const float* const*, const float* const*,
                    const int32_t*, int, float*, cudaStream_t) {
  set_error("tcgen05 edge path not built yet");
  return PG_ERR_UNSUPPORTED;
}

}  // namespace pg
