// Streaming (HBM-bound) pieces: activation fwd/bwd, MSE loss + its gradient, fused Adam.
//   loss/grad/Adam replace the user-side step of README.md:74-86
//   (jnp.mean((vmap(model)(x)-y)**2), optax.adam(3e-4), eqx.apply_updates).
#include "common.cuh"

namespace pdeqx {

static inline int stream_blocks(int64_t nvec) {
  int64_t b = ceil_div(nvec, 256);
  const int64_t cap = 148 * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

__global__ void __launch_bounds__(256)
act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int act) {
  const int64_t nv = n / 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float4 t = __ldg((const float4*)x + v);
    ((float4*)y)[v] = make_float4(apply_act_rt(t.x, act), apply_act_rt(t.y, act), apply_act_rt(t.z, act),
                                  apply_act_rt(t.w, act));
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = apply_act_rt(x[i], act);
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ gy, float* __restrict__ gx, int64_t n, int act) {
  const int64_t nv = n / 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float4 p = ((const float4*)pre)[v];  // may alias gx: plain loads, each element read before written
    float4 g = __ldg((const float4*)gy + v);
    ((float4*)gx)[v] = make_float4(g.x * act_grad_rt(p.x, act), g.y * act_grad_rt(p.y, act),
                                   g.z * act_grad_rt(p.z, act), g.w * act_grad_rt(p.w, act));
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    gx[i] = gy[i] * act_grad_rt(pre[i], act);
}

__global__ void __launch_bounds__(256)
add_kernel(const float* a, const float* b, float* out, int64_t n) {
  const int64_t nv = n / 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float4 p = ((const float4*)a)[v], q = ((const float4*)b)[v];  // out may alias a or b
    ((float4*)out)[v] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

constexpr int kMsePartials = 1024;

__global__ void __launch_bounds__(256)
mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ gpred,
                   float* __restrict__ partials, int64_t n, float gscale) {
  pdl_enter();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = pred[i] - target[i];
    acc = fmaf(d, d, acc);
    if (gpred) gpred[i] = gscale * d;
  }
  __shared__ float red[8];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffu, v, off);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256)
mse_final_kernel(const float* __restrict__ partials, int nparts, double inv_n, float* __restrict__ loss) {
  pdl_enter();
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += (double)partials[i];
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(red[0] * inv_n);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu, float* __restrict__ nu,
            int64_t n, float lr, float b1, float b2, float eps, float inv_c1, float inv_c2, float gscale,
            const float* __restrict__ corr) {
  pdl_enter();
  if (corr) {  // bias corrections of the device-side step counter (CUDA-graph replays)
    inv_c1 = corr[1];
    inv_c2 = corr[2];
  }
  const int64_t nv = n / 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = ((float4*)p)[v], gv = __ldg((const float4*)g + v), mv = ((float4*)mu)[v], vv = ((float4*)nu)[v];
    float pe[4] = {pv.x, pv.y, pv.z, pv.w}, ge[4] = {gv.x, gv.y, gv.z, gv.w};
    float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gg = ge[e] * gscale;
      me[e] = b1 * me[e] + (1.f - b1) * gg;
      ve[e] = b2 * ve[e] + (1.f - b2) * gg * gg;
      float upd = (me[e] * inv_c1) / (sqrtf(ve[e] * inv_c2) + eps);
      pe[e] -= lr * upd;
    }
    ((float4*)p)[v] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    ((float4*)mu)[v] = make_float4(me[0], me[1], me[2], me[3]);
    ((float4*)nu)[v] = make_float4(ve[0], ve[1], ve[2], ve[3]);
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gg = g[i] * gscale;
    float m = b1 * mu[i] + (1.f - b1) * gg;
    float v2 = b2 * nu[i] + (1.f - b2) * gg * gg;
    mu[i] = m;
    nu[i] = v2;
    p[i] -= lr * (m * inv_c1) / (sqrtf(v2 * inv_c2) + eps);
  }
}

}  // namespace pdeqx

using namespace pdeqx;

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int pdeqx_activation_fwd(int64_t n, const float* x, int32_t act, float* y, void* stream) {
  PDEQX_CHECK_ARG(n >= 0 && (n == 0 || (x && y)), "bad arguments");
  PDEQX_CHECK_ARG(act >= 0 && act <= 2, "unknown activation %d", act);
  PDEQX_CHECK_ARG(aligned16(x) && aligned16(y), "pointers must be 16-byte aligned");
  if (n == 0) return PDEQX_OK;
  act_fwd_kernel<<<stream_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(x, y, n, act);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_add(int64_t n, const float* a, const float* b, float* out, void* stream) {
  PDEQX_CHECK_ARG(n >= 0 && (n == 0 || (a && b && out)), "bad arguments");
  PDEQX_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(out), "pointers must be 16-byte aligned");
  if (n == 0) return PDEQX_OK;
  add_kernel<<<stream_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_activation_bwd(int64_t n, const float* pre, const float* gy, int32_t act, float* gx, void* stream) {
  PDEQX_CHECK_ARG(n >= 0 && (n == 0 || (pre && gy && gx)), "bad arguments");
  PDEQX_CHECK_ARG(act >= 0 && act <= 2, "unknown activation %d", act);
  PDEQX_CHECK_ARG(aligned16(pre) && aligned16(gy) && aligned16(gx), "pointers must be 16-byte aligned");
  if (n == 0) return PDEQX_OK;
  act_bwd_kernel<<<stream_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(pre, gy, gx, n, act);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_mse_loss(int64_t n, const float* pred, const float* target, float grad_scale, float* loss,
                              float* gpred, float* partials, void* stream) {
  PDEQX_CHECK_ARG(n > 0 && pred && target && loss && partials, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = (int)(ceil_div(n, 256 * 8) < kMsePartials ? ceil_div(n, 256 * 8) : kMsePartials);
  PDEQX_CUDA(launch_pdl(mse_partial_kernel, dim3(blocks), dim3(256), 0, s, pred, target, gpred, partials, n, grad_scale * 2.0f / (float)n));
  PDEQX_LAUNCHED();
  PDEQX_CUDA(launch_pdl(mse_final_kernel, dim3(1), dim3(256), 0, s, partials, blocks, 1.0 / (double)n, loss));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_adam_step(int64_t n, float* param, const float* grad, float* mu, float* nu, float lr, float b1,
                               float b2, float eps, int32_t step, float grad_scale, void* stream) {
  PDEQX_CHECK_ARG(n >= 0 && (n == 0 || (param && grad && mu && nu)), "bad arguments");
  PDEQX_CHECK_ARG(step >= 1, "step is the 1-based update count, got %d", step);
  PDEQX_CHECK_ARG(aligned16(param) && aligned16(grad) && aligned16(mu) && aligned16(nu),
                  "arena pointers must be 16-byte aligned");
  if (n == 0) return PDEQX_OK;
  const double c1 = 1.0 - pow((double)b1, (double)step), c2 = 1.0 - pow((double)b2, (double)step);
  adam_kernel<<<stream_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(param, grad, mu, nu, n, lr, b1, b2, eps,
                                                                         (float)(1.0 / c1), (float)(1.0 / c2), grad_scale, nullptr);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

// state[0] = update count (int32 bits), state[1] = 1 / (1 - b1^t), state[2] = 1 / (1 - b2^t)
__global__ void adam_tick_kernel(float* state, float b1, float b2) {
  pdl_enter();
  const int t = __float_as_int(state[0]) + 1;
  state[0] = __int_as_float(t);
  state[1] = (float)(1.0 / (1.0 - pow((double)b1, (double)t)));
  state[2] = (float)(1.0 / (1.0 - pow((double)b2, (double)t)));
}

extern "C" int pdeqx_adam_step_dev(int64_t n, float* param, const float* grad, float* mu, float* nu, float lr, float b1,
                                   float b2, float eps, float* state, float grad_scale, void* stream) {
  PDEQX_CHECK_ARG(n >= 0 && state && (n == 0 || (param && grad && mu && nu)), "bad arguments");
  PDEQX_CHECK_ARG(aligned16(param) && aligned16(grad) && aligned16(mu) && aligned16(nu),
                  "arena pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  PDEQX_CUDA(launch_pdl(adam_tick_kernel, dim3(1), dim3(1), 0, s, state, b1, b2));
  PDEQX_LAUNCHED();
  if (n == 0) return PDEQX_OK;
  PDEQX_CUDA(launch_pdl(adam_kernel, dim3(stream_blocks(n / 4 + 1)), dim3(256), 0, s, param, grad, mu, nu, n, lr, b1, b2, eps, 1.f, 1.f, grad_scale,
                        (const float*)state));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}
