// tcgen05 implicit-GEMM path of PhysicsConv (placeholder: not yet enabled; conv.cu's fp32 path runs).
#include "conv_plan.h"

namespace pdeqx {
bool tc_conv_available(const ConvGeom&) { return false; }
int tc_conv_fwd(const ConvGeom&, const float*, const float*, const float*, const float*, int, float*, cudaStream_t) {
  set_error("tc_conv_fwd: not available");
  return PDEQX_ERR_UNSUPPORTED;
}
int tc_conv_bwd_data(const ConvGeom&, const float*, const float*, const float*, float*, cudaStream_t) {
  set_error("tc_conv_bwd_data: not available");
  return PDEQX_ERR_UNSUPPORTED;
}
int tc_conv_bwd_filter(const ConvGeom&, const float*, const float*, float*, float*, cudaStream_t) {
  set_error("tc_conv_bwd_filter: not available");
  return PDEQX_ERR_UNSUPPORTED;
}
}  // namespace pdeqx
