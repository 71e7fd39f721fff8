// Library-wide state: error string, launch counter, fast-path switch, version.
#include <string.h>

#include "common.cuh"

namespace pdeqx {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_fast_paths{1};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace pdeqx

extern "C" const char* pdeqx_last_error(void) { return pdeqx::t_error; }
extern "C" const char* pdeqx_version(void) { return "0.1.0 (sm_100a)"; }
extern "C" uint64_t pdeqx_launch_count(void) { return pdeqx::g_launches.load(std::memory_order_relaxed); }

extern "C" int pdeqx_has_tcgen05(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" void pdeqx_set_fast_paths(int enabled) { pdeqx::g_fast_paths.store(enabled ? 1 : 0); }
