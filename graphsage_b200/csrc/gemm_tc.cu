// tcgen05 path of gs_sage_gemm - placeholder until the UMMA kernel lands (next milestone).
#include "common.cuh"

namespace gs {
int64_t sage_gemm_tc_workspace(int64_t, const gs_gemm_part*, int32_t, int32_t) { return 0; }
int32_t sage_gemm_tc(int64_t, const gs_gemm_part*, int32_t, int32_t, const float*, int32_t, int32_t math, float*, int64_t,
                     void*, cudaStream_t) {
  set_error("gs_sage_gemm: math mode %d (tcgen05) not built yet", math);
  return GS_ERR_UNSUPPORTED;
}
}  // namespace gs
