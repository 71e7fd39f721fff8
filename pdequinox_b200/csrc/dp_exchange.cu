// The one exchange step of the batch-sharded train step as ONE kernel over NVLink peer memory:
//
//     reduce-scatter of the gradient arenas  +  Adam on the rank's 1/N shard  +  all-gather of the updated parameters
//
// instead of ncclAllReduce (67 MB, ~0.33 ms with its stream joins at 8 ranks) followed by a replicated Adam pass over the
// whole arena (68 us). Every rank owns the shard [rank * n / N, (rank + 1) * n / N): it LOADS that slice of all N gradient
// arenas straight from the peers' memory (NVSwitch gives every pair full bandwidth), sums them in rank order, applies the
// Adam update (moments exist for the shard only: 1/N of the optimiser memory) and STORES the new parameters into all N
// parameter arenas. Per rank 7/8 of 67 MB cross NVLink in each direction once -- half of what a ring all-reduce moves --
// and no SM is held by a communication library while the compute kernels run. The scalar loss travels the same way.
//
// Ordering between ranks uses two flag barriers in peer memory (system-scope release / acquire):
//   A  "my gradients and loss are final"   -- entered by the kernel itself before it reads the peers' gradients;
//   B  "my parameter stores are complete"  -- a second, one-CTA kernel behind the first (the kernel boundary orders every
//      store of the first kernel before the flag), left when all ranks have published: the next forward pass may start.
// Both kernels are ordinary stream work: the whole step still replays as one CUDA graph. A wait that outlasts
// PDEQX_DP_TIMEOUT cycles (a peer died) raises an error word in the rank's own signal pad instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace pdeqx {

constexpr int kPadFlagsA = 0, kPadFlagsB = PDEQX_MAX_RANKS, kPadEpoch = 2 * PDEQX_MAX_RANKS, kPadError = 2 * PDEQX_MAX_RANKS + 1;
constexpr int kPadLoss = 32;  // float slots [PDEQX_MAX_RANKS]
constexpr long long kDpTimeoutCycles = 6000000000ll;  // ~3 s at 1.9 GHz

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// NVSwitch multicast (NVLS): one load that returns the SUM over every device bound to the multicast object / one store that
// the switch replicates to all of them
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// spin until *flag >= epoch (wrap-safe); false on timeout
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (clock64() - t0 > kDpTimeoutCycles) return false;
    __nanosleep(64);
  }
  return true;
}

__global__ void __launch_bounds__(256)
dp_adam_kernel(pdeqx_dp_peers P, int64_t n, float* __restrict__ mu, float* __restrict__ nu, float lr, float b1, float b2, float eps,
               const float* __restrict__ corr, float* __restrict__ loss_io) {
  uint32_t* mypad = P.pads[P.rank];
  const uint32_t epoch = *(volatile uint32_t*)(mypad + kPadEpoch) + 1;
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();
  // ---- barrier A: publish "my gradients / loss are final" (block 0), wait for everybody (every block) ----
  if (blockIdx.x == 0 && threadIdx.x < P.world) {
    const int r = threadIdx.x;
    if (loss_io) ((volatile float*)(P.pads[r] + kPadLoss))[P.rank] = *loss_io;
    __threadfence_system();
    st_release_sys(P.pads[r] + kPadFlagsA + P.rank, epoch);
  }
  if (threadIdx.x < P.world && !wait_flag(mypad + kPadFlagsA + threadIdx.x, epoch)) {
    ok_s = 0;
    mypad[kPadError] = 1;
  }
  __syncthreads();
  if (!ok_s) return;
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_io) {
    float s = 0.f;
    for (int r = 0; r < P.world; ++r) s += ((volatile float*)(mypad + kPadLoss))[r];
    *loss_io = s / (float)P.world;
  }
  // ---- the rank's shard: sum of the N gradient slices -> Adam -> new parameters into all N arenas ----
  const float inv_c1 = corr[1], inv_c2 = corr[2], gscale = 1.f / (float)P.world;
  const int64_t nv = n / 4, per = (nv + P.world - 1) / P.world;
  const int64_t lo = (int64_t)P.rank * per, hi = lo + per < nv ? lo + per : nv;
  float4* const own = (float4*)P.params[P.rank];
  for (int64_t v = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < hi; v += (int64_t)gridDim.x * blockDim.x) {
    // all N loads are issued before the first add (a peer load takes microseconds: serialised they would bound the kernel);
    // peer memory is rewritten every step: ld.cv, never from a stale cache line
    float4 gs;
    if (P.mc_grads) {
      gs = multimem_ld_reduce_add((const float4*)P.mc_grads + v);  // summed inside the switch
    } else {
      float4 g[PDEQX_MAX_RANKS];
#pragma unroll
      for (int r = 0; r < PDEQX_MAX_RANKS; ++r)
        g[r] = r < P.world ? __ldcv((const float4*)P.grads[r] + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      gs = g[0];
#pragma unroll
      for (int r = 1; r < PDEQX_MAX_RANKS; ++r) {  // fixed order: every replica of the job computes bit-identical parameters
        gs.x += g[r].x; gs.y += g[r].y; gs.z += g[r].z; gs.w += g[r].w;
      }
    }
    const float4 pv = own[v], mv = ((float4*)mu)[v - lo], vv = ((float4*)nu)[v - lo];
    float pe[4] = {pv.x, pv.y, pv.z, pv.w}, ge[4] = {gs.x, gs.y, gs.z, gs.w};
    float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = ge[e] * gscale;
      me[e] = b1 * me[e] + (1.f - b1) * gg;
      ve[e] = b2 * ve[e] + (1.f - b2) * gg * gg;
      pe[e] -= lr * ((me[e] * inv_c1) / (sqrtf(ve[e] * inv_c2) + eps));
    }
    ((float4*)mu)[v - lo] = make_float4(me[0], me[1], me[2], me[3]);
    ((float4*)nu)[v - lo] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    const float4 pn = make_float4(pe[0], pe[1], pe[2], pe[3]);
    if (P.mc_params) {
      multimem_st((float4*)P.mc_params + v, pn);  // replicated by the switch into every rank's arena (this one included)
    } else {
      for (int r = 0; r < P.world; ++r) ((float4*)P.params[r])[v] = pn;
    }
  }
}

// barrier B: all parameter stores of this rank (previous kernel) are complete -> tell everybody, wait for everybody
__global__ void dp_publish_kernel(pdeqx_dp_peers P) {
  uint32_t* mypad = P.pads[P.rank];
  const uint32_t epoch = *(volatile uint32_t*)(mypad + kPadEpoch) + 1;
  __threadfence_system();
  bool ok = true;
  if (threadIdx.x < P.world) {
    st_release_sys(P.pads[threadIdx.x] + kPadFlagsB + P.rank, epoch);
    ok = wait_flag(mypad + kPadFlagsB + threadIdx.x, epoch);
    if (!ok) mypad[kPadError] = 1;
  }
  __syncwarp();
  if (threadIdx.x == 0) *(volatile uint32_t*)(mypad + kPadEpoch) = epoch;
}

// state[0] = update count (int32 bits), state[1] = 1 / (1 - b1^t), state[2] = 1 / (1 - b2^t)   (as pdeqx_adam_step_dev)
__global__ void dp_tick_kernel(float* state, float b1, float b2) {
  const int t = __float_as_int(state[0]) + 1;
  state[0] = __int_as_float(t);
  state[1] = (float)(1.0 / (1.0 - pow((double)b1, (double)t)));
  state[2] = (float)(1.0 / (1.0 - pow((double)b2, (double)t)));
}

}  // namespace pdeqx

using namespace pdeqx;

extern "C" int pdeqx_dp_alloc(size_t bytes, void** out) {
  PDEQX_CHECK_ARG(out && bytes > 0, "bad arguments");
  PDEQX_CUDA(cudaMalloc(out, bytes));  // a plain allocation of its own: exportable with cudaIpcGetMemHandle (base pointer)
  PDEQX_CUDA(cudaMemset(*out, 0, bytes));
  return PDEQX_OK;
}
extern "C" int pdeqx_dp_free(void* p) {
  PDEQX_CUDA(cudaFree(p));
  return PDEQX_OK;
}
extern "C" int pdeqx_dp_export(const void* p, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  PDEQX_CHECK_ARG(p && handle, "null argument");
  cudaIpcMemHandle_t h;
  PDEQX_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
  memcpy(handle, &h, 64);
  return PDEQX_OK;
}
extern "C" int pdeqx_dp_import(const uint8_t handle[64], void** out) {
  PDEQX_CHECK_ARG(out && handle, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  PDEQX_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return PDEQX_OK;
}
extern "C" int pdeqx_dp_unimport(void* p) {
  PDEQX_CUDA(cudaIpcCloseMemHandle(p));
  return PDEQX_OK;
}
extern "C" size_t pdeqx_dp_pad_bytes(void) { return 64 * sizeof(uint32_t); }
extern "C" int64_t pdeqx_dp_shard_floats(int64_t n, int32_t world) { return world > 0 ? ((n / 4 + world - 1) / world) * 4 : 0; }

extern "C" int pdeqx_dp_adam_step(const pdeqx_dp_peers* peers, int64_t n, float* mu_shard, float* nu_shard, float lr, float b1,
                                  float b2, float eps, float* state, float* loss_io, void* stream) {
  PDEQX_CHECK_ARG(peers && state && mu_shard && nu_shard, "null argument");
  PDEQX_CHECK_ARG(peers->world >= 1 && peers->world <= PDEQX_MAX_RANKS && peers->rank >= 0 && peers->rank < peers->world,
                  "world / rank out of range (at most %d ranks)", PDEQX_MAX_RANKS);
  PDEQX_CHECK_ARG(n >= 0 && n % 4 == 0, "the arena length must be a multiple of 4 floats");
  for (int r = 0; r < peers->world; ++r)
    PDEQX_CHECK_ARG(peers->grads[r] && peers->params[r] && peers->pads[r] && ((uintptr_t)peers->grads[r] & 15) == 0 &&
                        ((uintptr_t)peers->params[r] & 15) == 0,
                    "peer %d: null or unaligned arena pointers", r);
  cudaStream_t s = (cudaStream_t)stream;
  TimedScope timed(PDEQX_K_DP_EXCHANGE, s);
  // bytes over this GPU's links (in + out) + the local p / mu / nu traffic of the shard
  timed_bytes(4.0 * (double)n * ((peers->mc_grads ? 2.0 : 2.0 * (peers->world - 1)) / peers->world + 4.0 / peers->world));
  dp_tick_kernel<<<1, 1, 0, s>>>(state, b1, b2);
  PDEQX_LAUNCHED();
  const int64_t per = pdeqx_dp_shard_floats(n, peers->world) / 4;
  int64_t blocks = ceil_div(per > 0 ? per : 1, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dp_adam_kernel<<<(int)blocks, 256, 0, s>>>(*peers, n, mu_shard, nu_shard, lr, b1, b2, eps, state, loss_io);
  PDEQX_LAUNCHED();
  dp_publish_kernel<<<1, 32, 0, s>>>(*peers);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

// error word of the rank's own signal pad (a peer did not arrive within the timeout); synchronises the stream
extern "C" int pdeqx_dp_check(const pdeqx_dp_peers* peers, void* stream) {
  PDEQX_CHECK_ARG(peers, "null argument");
  uint32_t err = 0;
  PDEQX_CUDA(cudaMemcpyAsync(&err, peers->pads[peers->rank] + kPadError, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  PDEQX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (err) {
    set_error("data-parallel exchange: a peer rank did not arrive within the timeout");
    return PDEQX_ERR_CUDA;
  }
  return PDEQX_OK;
}
