// Shared helpers for libpdeqx_b200 (sm_100a only).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>

#include "../../include/pdeqx_b200.h"

namespace pdeqx {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_fast_paths;
// ---- programmatic dependent launch (experiment, PDEQX_PDL=1 switches it on; see pdl_enabled() in core.cu) ------
// With the switch on every kernel of the spectral train step is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts
// with pdl_enter(): griddepcontrol.launch_dependents (the NEXT kernel's CTAs may be scheduled as soon as every CTA of this
// grid has started) followed by griddepcontrol.wait (nothing of global memory is touched, and no exclusive SM resource such
// as tensor memory is taken, before the previous grid has completed and flushed). What overlaps is the launch latency and
// the CTA ramp-up of kernel N + 1 with the tail of kernel N -- a replayed step is ~70 kernels, a dozen of them 2 - 8 us long.
bool pdl_enabled();
template <typename... Params, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// a[0:na] = b[0:nb] = 0 by one small kernel (either pointer may be null): see core.cu
int zero_small(float* a, int64_t na, float* b, int64_t nb, cudaStream_t s);

#define PDEQX_CHECK_ARG(cond, ...)                 \
  do {                                             \
    if (!(cond)) {                                 \
      ::pdeqx::set_error(__VA_ARGS__);             \
      return PDEQX_ERR_INVALID_ARGUMENT;           \
    }                                              \
  } while (0)

#define PDEQX_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::pdeqx::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                         __FILE__, __LINE__);                                         \
      return PDEQX_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define PDEQX_LAUNCHED()                                                              \
  do {                                                                                \
    ::pdeqx::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    cudaError_t e__ = cudaPeekAtLastError();                                          \
    if (e__ != cudaSuccess) {                                                         \
      ::pdeqx::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                         \
      return PDEQX_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define PDEQX_TRY(call)        \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != PDEQX_OK) return rc__; \
  } while (0)

// ---- optional per-kernel timing (core.cu) --------------------------------------------------
struct TimedScope {
  int tag;
  cudaStream_t stream;
  int slot;
  double bytes;       // algorithmic bytes of the launches inside the scope (reported by the launchers, see timed_bytes)
  TimedScope* outer;  // enclosing scope of the calling thread
  TimedScope(int tag, cudaStream_t s);
  ~TimedScope();
};
// A launcher states the ALGORITHMIC bytes of the launch it just enqueued (the tensors it reads and writes once each --
// SURVEY.md 8d -- not what the hardware moved): credited to the innermost timed scope of the calling thread, no-op
// when timing is off. bench.py divides these by the scope's event time for its roofline line.
void timed_bytes(double bytes);

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Function attributes (the 227 KB dynamic shared-memory opt-in) are per DEVICE: a launcher keeps one of these per kernel
// and configures the kernel on the first launch on every device of the process.
struct DeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool first_use() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const uint64_t bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
  }
};

// ---- activations (device) -------------------------------------------------
// tanh via exp: |err| ~ 1e-7 (fp32 parity mode needs better than tanh.approx's 2^-11)
__device__ __forceinline__ float tanh_accurate(float u) {
  float a = fabsf(u);
  float e = __expf(-2.0f * a);            // in (0,1]
  float t = __fdividef(1.0f - e, 1.0f + e);
  return copysignf(t, u);
}

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanh_accurate(u));
}

__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float u = k0 * (x + k1 * x * x2);
  float t = tanh_accurate(u);
  float du = k0 * (1.0f + 3.0f * k1 * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
  if (ACT == PDEQX_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == PDEQX_ACT_GELU_TANH) return gelu_tanh(v);
  return v;
}

__device__ __forceinline__ float apply_act_rt(float v, int act) {
  if (act == PDEQX_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == PDEQX_ACT_GELU_TANH) return gelu_tanh(v);
  return v;
}

__device__ __forceinline__ float act_grad_rt(float pre, int act) {
  if (act == PDEQX_ACT_RELU) return pre > 0.0f ? 1.0f : 0.0f;
  if (act == PDEQX_ACT_GELU_TANH) return gelu_tanh_grad(pre);
  return 1.0f;
}

// ---- generic GEMM front-ends (gemm_generic.cu) ------------------------------
// C[b] (M x N, row-major, ldc) = epilogue( A[b] (M x K) . B[b] (K x N) )
//   A element (m,k) at A + b*strideA + m*lda_m + k*lda_k  (so A or A^T)
//   B element (k,n) at B + b*strideB + k*ldb + n          (n contiguous)
//   epilogue: v = acc + (bias ? bias[m] : 0) + (add ? add[b*strideC + m*ldc + n] : 0); C = act(v)
struct SgemmArgs {
  const float* A; const float* B; float* C;
  const float* bias; const float* add;
  int64_t M, N, K;
  int64_t lda_m, lda_k, ldb, ldc;
  int64_t strideA, strideB, strideC;
  int32_t batch;
  int32_t act;
};
int sgemm_batched(const SgemmArgs& a, cudaStream_t s);

// Complex batched GEMM with a shared complex A (twiddle table):
//   C[b][m][n] = sum_k A[m][k] * B[b][k][n], complex interleaved (float2), n contiguous.
int cgemm_shared_a(const float2* A, const float2* B, float2* C, int64_t batch, int M, int K, int64_t N,
                   cudaStream_t s);

}  // namespace pdeqx
