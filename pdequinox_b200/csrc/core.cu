// Library-wide state: error string, launch counter, fast-path switch, version.
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace pdeqx {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_fast_paths{1};

// ---- small accumulators zeroed by a KERNEL node -------------------------------------------------
// Inside a replayed CUDA graph a memset node costs ~3 us of idle time on top of its own 1.5 us (CUPTI timeline of the train step,
// profiles/r2_graph_timeline_b8.txt), a kernel node none: the per-block gradient accumulators (a few KB each, 13 per step) are
// cleared by this kernel, two buffers per launch.
__global__ void zero2_kernel(float* __restrict__ a, int64_t na, float* __restrict__ b, int64_t nb) {
  pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < na) a[i] = 0.f;
    else b[i - na] = 0.f;
  }
}
bool pdl_enabled() {
  // off unless PDEQX_PDL=1: measured SLOWER on B200 (configs[3] step 6.269 -> 6.365 ms, 1.199 -> 1.228 ms at 8 samples per GPU):
  // the CUPTI timeline of the replayed graph shows no idle time between kernel nodes to win back, and the early-scheduled CTAs
  // of kernel N + 1 hold SM slots while they wait
  static const bool on = [] { const char* e = getenv("PDEQX_PDL"); return e && e[0] == '1'; }();
  return on;
}

int zero_small(float* a, int64_t na, float* b, int64_t nb, cudaStream_t s) {
  if (!a) na = 0;
  if (!b) nb = 0;
  if (na + nb == 0) return PDEQX_OK;
  const int64_t blocks = (na + nb + 255) / 256;
  PDEQX_CUDA(launch_pdl(zero2_kernel, dim3((unsigned)(blocks < 64 ? blocks : 64)), dim3(256), 0, s, a, na, b, nb));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

// ---- per-kernel timing ---------------------------------------------------------------------
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
struct TimedRec { cudaEvent_t a, b; int tag; double bytes; };
static thread_local TimedScope* t_scope = nullptr;
static std::vector<TimedRec> g_timed;
constexpr size_t kMaxTimed = 1 << 16;

TimedScope::TimedScope(int t, cudaStream_t s) : tag(t), stream(s), slot(-1), bytes(0.0), outer(t_scope) {
  t_scope = this;
  if (!g_timing.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (g_timed.size() >= kMaxTimed) return;
  TimedRec r{};
  r.tag = t;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  slot = (int)g_timed.size();
  g_timed.push_back(r);
}

TimedScope::~TimedScope() {
  t_scope = outer;
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if ((size_t)slot < g_timed.size()) {
    cudaEventRecord(g_timed[slot].b, stream);
    g_timed[slot].bytes = bytes;
  }
}

void timed_bytes(double b) {
  if (t_scope && t_scope->slot >= 0) t_scope->bytes += b;
}

}  // namespace pdeqx

extern "C" void pdeqx_timing_enable(int enabled) {
  using namespace pdeqx;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& r : g_timed) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_timed.clear();
  g_timing.store(enabled ? 1 : 0);
}

extern "C" int pdeqx_timing_read_ex(int32_t tag, double* total_ms, int64_t* launches, double* algorithmic_bytes) {
  using namespace pdeqx;
  PDEQX_CHECK_ARG(total_ms && launches, "null argument");
  std::lock_guard<std::mutex> lk(g_timing_mu);
  double tot = 0.0, bytes = 0.0;
  int64_t n = 0;
  for (auto& r : g_timed) {
    if (r.tag != tag) continue;
    PDEQX_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    PDEQX_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    tot += ms;
    bytes += r.bytes;
    ++n;
  }
  *total_ms = tot;
  *launches = n;
  if (algorithmic_bytes) *algorithmic_bytes = bytes;
  return PDEQX_OK;
}

extern "C" int pdeqx_timing_read(int32_t tag, double* total_ms, int64_t* launches) {
  return pdeqx_timing_read_ex(tag, total_ms, launches, nullptr);
}

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

extern "C" int pdeqx_current_device(void) {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

extern "C" void pdeqx_set_fast_paths(int enabled) { pdeqx::g_fast_paths.store(enabled ? 1 : 0); }
