// tcgen05 stages of the spectral block (placeholder: not yet enabled; the fp32 GEMM stages run).
#include "spectral_plan.h"

namespace pdeqx {
int tc_plan_init(pdeqx_spectral_plan*) { return PDEQX_OK; }
void tc_plan_free(pdeqx_spectral_plan*) {}
bool tc_r2c_available(const pdeqx_spectral_plan*) { return false; }
int tc_r2c(const pdeqx_spectral_plan*, const float*, bool, int64_t, float*, cudaStream_t) { return PDEQX_ERR_UNSUPPORTED; }
bool tc_slab_available(const pdeqx_spectral_plan*, bool) { return false; }
int tc_slab(const pdeqx_spectral_plan*, const float*, const float*, int, const float*, bool, const float*, int, float*,
            cudaStream_t) {
  return PDEQX_ERR_UNSUPPORTED;
}
}  // namespace pdeqx
