// PointwiseLinearConv (1x1 conv) forward / backward.
//
//   replaces pdequinox/conv/_pointwise_linear_conv.py:6-53 (eqx.nn.Conv with kernel_size=1):
//   y[b,o,n] = sum_i w[o,i] x[b,i,n] + bias[o]
//
// Three regimes:
//   * lifting (c_in tiny) / projection (c_out tiny): streaming kernels, one pass over the big
//     tensor, float4 loads/stores -- pure HBM-bound (0.5 flop/B);
//   * c_in, c_out both large: GEMM (CUDA-core fp32 here; the FNO block's bypass runs fused into
//     the tcgen05 slab kernel instead, see tc_spectral.cu);
//   * weight gradient: persistent split-K GEMM over all pixels with one atomic flush per CTA.
#include "common.cuh"

namespace pdeqx {

constexpr int kSmallC = 8;

// ---- few input channels: keep x in registers, loop over outputs ----------------------------
//  W(o,i) = w[o*w_so + i*w_si]  (lets the backward pass read w transposed)
template <int VEC>
__global__ void __launch_bounds__(256)
pw_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    const float* __restrict__ add, float* __restrict__ y, int cin, int cout, int64_t n,
                    int w_so, int w_si, int act) {
  extern __shared__ float sw[];  // [cout][cin] + [cout]
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) sw[i] = w[(i / cin) * w_so + (i % cin) * w_si];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[cin * cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t nv = n / VEC;
  const float* xb = x + (int64_t)b * cin * n;
  float* yb = y + (int64_t)b * cout * n;
  const float* ab = add ? add + (int64_t)b * cout * n : nullptr;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float xr[kSmallC][VEC];
    for (int i = 0; i < cin; ++i) {
      if (VEC == 4) {
        float4 t = __ldg((const float4*)(xb + (int64_t)i * n) + v);
        xr[i][0] = t.x; xr[i][1 % VEC] = t.y; xr[i][2 % VEC] = t.z; xr[i][3 % VEC] = t.w;
      } else {
        xr[i][0] = __ldg(xb + (int64_t)i * n + v);
      }
    }
    for (int o = 0; o < cout; ++o) {
      float acc[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] = sw[cin * cout + o];
      for (int i = 0; i < cin; ++i) {
        float wv = sw[o * cin + i];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(wv, xr[i][e], acc[e]);
      }
      if (VEC == 4) {
        if (ab) {
          float4 t = __ldg((const float4*)(ab + (int64_t)o * n) + v);
          acc[0] += t.x; acc[1 % VEC] += t.y; acc[2 % VEC] += t.z; acc[3 % VEC] += t.w;
        }
        float4 r = make_float4(apply_act_rt(acc[0], act), apply_act_rt(acc[1 % VEC], act),
                               apply_act_rt(acc[2 % VEC], act), apply_act_rt(acc[3 % VEC], act));
        ((float4*)(yb + (int64_t)o * n))[v] = r;
      } else {
        if (ab) acc[0] += __ldg(ab + (int64_t)o * n + v);
        yb[(int64_t)o * n + v] = apply_act_rt(acc[0], act);
      }
    }
  }
}

// ---- few output channels: accumulate in registers, stream the inputs -----------------------
template <int VEC>
__global__ void __launch_bounds__(256)
pw_small_cout_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ add, float* __restrict__ y, int cin, int cout, int64_t n,
                     int w_so, int w_si, int act) {
  extern __shared__ float sw[];  // [cout][cin] + [cout]
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) sw[i] = w[(i / cin) * w_so + (i % cin) * w_si];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[cin * cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t nv = n / VEC;
  const float* xb = x + (int64_t)b * cin * n;
  float* yb = y + (int64_t)b * cout * n;
  const float* ab = add ? add + (int64_t)b * cout * n : nullptr;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float acc[kSmallC][VEC];
#pragma unroll
    for (int o = 0; o < kSmallC; ++o)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[o][e] = o < cout ? sw[cin * cout + o] : 0.f;
    // 16 channel rows in flight per thread: the kernel is a pure stream over x (one 16-byte load per channel and thread)
    for (int i0 = 0; i0 < cin; i0 += 16) {
      float xv[16][VEC];
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) {
        const int i = i0 + ii < cin ? i0 + ii : cin - 1;
        if (VEC == 4) {
          float4 t = __ldg((const float4*)(xb + (int64_t)i * n) + v);
          xv[ii][0] = t.x; xv[ii][1 % VEC] = t.y; xv[ii][2 % VEC] = t.z; xv[ii][3 % VEC] = t.w;
        } else {
          xv[ii][0] = __ldg(xb + (int64_t)i * n + v);
        }
      }
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) {
        if (i0 + ii < cin) {
#pragma unroll
          for (int o = 0; o < kSmallC; ++o) {  // static register indices: `cout` only predicates
            if (o < cout) {
              float wv = sw[o * cin + i0 + ii];
#pragma unroll
              for (int e = 0; e < VEC; ++e) acc[o][e] = fmaf(wv, xv[ii][e], acc[o][e]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < kSmallC; ++o) {
      if (o >= cout) break;
      if (VEC == 4) {
        if (ab) {
          float4 t = __ldg((const float4*)(ab + (int64_t)o * n) + v);
          acc[o][0] += t.x; acc[o][1 % VEC] += t.y; acc[o][2 % VEC] += t.z; acc[o][3 % VEC] += t.w;
        }
        ((float4*)(yb + (int64_t)o * n))[v] =
            make_float4(apply_act_rt(acc[o][0], act), apply_act_rt(acc[o][1 % VEC], act),
                        apply_act_rt(acc[o][2 % VEC], act), apply_act_rt(acc[o][3 % VEC], act));
      } else {
        if (ab) acc[o][0] += __ldg(ab + (int64_t)o * n + v);
        yb[(int64_t)o * n + v] = apply_act_rt(acc[o][0], act);
      }
    }
  }
}

// generic W(o,i) strides so that bwd-data can reuse the forward kernels with w^T
static int pointwise_apply(int batch, int cin, int cout, int64_t n, const float* x, const float* w, int w_so, int w_si,
                           const float* bias, const float* add, int act, float* y, cudaStream_t s) {
  if (batch <= 0 || n <= 0) return PDEQX_OK;
  PDEQX_CHECK_ARG(batch <= 65535, "pointwise: batch %d exceeds grid.y", batch);
  const bool vec4 = (n % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                    (!add || (uintptr_t)add % 16 == 0);
  const int64_t nv = vec4 ? n / 4 : n;
  const int blocks = (int)(ceil_div(nv, 256) < 2048 ? ceil_div(nv, 256) : 2048);
  const size_t smem = (size_t)(cin * cout + cout) * sizeof(float);
  if (cin <= kSmallC && smem <= 48 * 1024) {
    if (vec4)
      pw_small_cin_kernel<4><<<dim3(blocks, batch), 256, smem, s>>>(x, w, bias, add, y, cin, cout, n, w_so, w_si, act);
    else
      pw_small_cin_kernel<1><<<dim3(blocks, batch), 256, smem, s>>>(x, w, bias, add, y, cin, cout, n, w_so, w_si, act);
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
  if (cout <= kSmallC && smem <= 48 * 1024) {
    if (vec4)
      pw_small_cout_kernel<4><<<dim3(blocks, batch), 256, smem, s>>>(x, w, bias, add, y, cin, cout, n, w_so, w_si, act);
    else
      pw_small_cout_kernel<1><<<dim3(blocks, batch), 256, smem, s>>>(x, w, bias, add, y, cin, cout, n, w_so, w_si, act);
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
  SgemmArgs a{};
  a.A = w; a.B = x; a.C = y; a.bias = bias; a.add = add;
  a.M = cout; a.N = n; a.K = cin;
  a.lda_m = w_so; a.lda_k = w_si; a.ldb = n; a.ldc = n;
  a.strideA = 0; a.strideB = (int64_t)cin * n; a.strideC = (int64_t)cout * n;
  a.batch = batch; a.act = act;
  return sgemm_batched(a, s);
}

// ---- weight / bias gradient ----------------------------------------------------------------
// General case: gw[64x64 tile] += G[o, K-chunk] . X[i, K-chunk]^T, persistent over (b, pixel chunks).
constexpr int WG_BK = 32;
__global__ void __launch_bounds__(256)
pw_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gw,
                float* __restrict__ gb, int batch, int cin, int cout, int64_t n, int64_t chunks_per_b) {
  __shared__ float Gs[WG_BK][64 + 4];
  __shared__ float Xs[WG_BK][64 + 4];
  const int tid = threadIdx.x;
  const int o0 = blockIdx.y * 64, i0 = blockIdx.z * 64;
  const int ty = tid / 16, tx = tid % 16;  // 4x4 micro tile
  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int64_t total = (int64_t)batch * chunks_per_b;
  for (int64_t c = blockIdx.x; c < total; c += gridDim.x) {
    const int b = (int)(c / chunks_per_b);
    const int64_t n0 = (c % chunks_per_b) * WG_BK;
    // load [64 ch][WG_BK px] tiles, transposed into [px][ch]
    for (int idx = tid; idx < 64 * WG_BK; idx += 256) {
      int ch = idx / WG_BK, px = idx % WG_BK;
      int64_t gn = n0 + px;
      float gv = 0.f, xv = 0.f;
      if (gn < n) {
        if (o0 + ch < cout) gv = __ldg(gy + ((int64_t)b * cout + o0 + ch) * n + gn);
        if (i0 + ch < cin) xv = __ldg(x + ((int64_t)b * cin + i0 + ch) * n + gn);
      }
      Gs[px][ch] = gv;
      Xs[px][ch] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Gs[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Xs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bsum[i] += a[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int o = o0 + ty * 4 + i;
    if (o >= cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ii = i0 + tx * 4 + j;
      if (ii < cin) atomicAdd(gw + (int64_t)o * cin + ii, acc[i][j]);
    }
    if (gb && tx == 0 && blockIdx.z == 0) atomicAdd(gb + o, bsum[i]);
  }
}

// Small side (<= kSmallC channels on one of the two sides; the lifting 1 -> C and projection C -> 1 of every
// architecture): persistent streaming kernel. A block walks (batch, pixel-chunk) units; the small side's chunk is
// staged in shared memory, each warp owns a strided set of channels of the big side and keeps their partial sums in
// registers across ALL its units (float4 loads, 4 in flight per lane), one atomic flush per block at the end.
constexpr int WS_CHUNK = 1024;     // pixels per unit
constexpr int WS_MAXCH = 16;       // big-side channels per warp (8 warps -> up to 128 channels)
__global__ void __launch_bounds__(256)
pw_wgrad_small_kernel(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gw,
                      float* __restrict__ gb, int batch, int cin, int cout, int64_t n) {
  extern __shared__ float sm_small[];  // [nsmall][WS_CHUNK]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool big_is_gy = cout >= cin;
  const int nbig = big_is_gy ? cout : cin, nsmall = big_is_gy ? cin : cout;
  const int64_t chunks = (n + WS_CHUNK - 1) / WS_CHUNK;
  const int64_t units = (int64_t)batch * chunks;
  float acc[WS_MAXCH][kSmallC];
  float bsum[WS_MAXCH];
  float ssum[kSmallC];  // bias gradient over the small side (only when the small side is gy)
#pragma unroll
  for (int c = 0; c < WS_MAXCH; ++c) {
    bsum[c] = 0.f;
#pragma unroll
    for (int j = 0; j < kSmallC; ++j) acc[c][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < kSmallC; ++j) ssum[j] = 0.f;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = (int)(u / chunks);
    const int64_t p0 = (u % chunks) * WS_CHUNK;
    const int len = (int)((n - p0) < WS_CHUNK ? (n - p0) : WS_CHUNK);
    const float* bigp = (big_is_gy ? gy + (int64_t)b * cout * n : x + (int64_t)b * cin * n) + p0;
    const float* smallp = (big_is_gy ? x + (int64_t)b * cin * n : gy + (int64_t)b * cout * n) + p0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < nsmall * WS_CHUNK; idx += blockDim.x) {
      const int j = idx / WS_CHUNK, q = idx % WS_CHUNK;
      sm_small[idx] = q < len ? __ldg(smallp + (int64_t)j * n + q) : 0.f;
    }
    __syncthreads();
    const bool vec = (len == WS_CHUNK) && ((n & 3) == 0) && (((uintptr_t)bigp & 15) == 0);
#pragma unroll
    for (int ci = 0; ci < WS_MAXCH; ++ci) {
      const int c = warp + ci * 8;
      if (c >= nbig) break;
      const float* row = bigp + (int64_t)c * n;
      if (vec) {
        float4 v[WS_CHUNK / 128];
#pragma unroll
        for (int t = 0; t < WS_CHUNK / 128; ++t) v[t] = __ldg((const float4*)row + lane + 32 * t);
#pragma unroll
        for (int t = 0; t < WS_CHUNK / 128; ++t) {
          bsum[ci] += (v[t].x + v[t].y) + (v[t].z + v[t].w);
#pragma unroll
          for (int j = 0; j < kSmallC; ++j) {
            if (j < nsmall) {
              const float4 sv = *(const float4*)(sm_small + j * WS_CHUNK + (lane + 32 * t) * 4);
              acc[ci][j] = fmaf(v[t].x, sv.x, fmaf(v[t].y, sv.y, fmaf(v[t].z, sv.z, fmaf(v[t].w, sv.w, acc[ci][j]))));
            }
          }
        }
      } else {
        for (int q = lane; q < len; q += 32) {
          const float v = __ldg(row + q);
          bsum[ci] += v;
#pragma unroll
          for (int j = 0; j < kSmallC; ++j)
            if (j < nsmall) acc[ci][j] = fmaf(v, sm_small[j * WS_CHUNK + q], acc[ci][j]);
        }
      }
    }
    if (gb && !big_is_gy && warp == 0) {
#pragma unroll
      for (int j = 0; j < kSmallC; ++j)
        if (j < nsmall)
          for (int q = lane; q < len; q += 32) ssum[j] += sm_small[j * WS_CHUNK + q];
    }
  }
  // flush
#pragma unroll
  for (int ci = 0; ci < WS_MAXCH; ++ci) {
    const int c = warp + ci * 8;
    if (c >= nbig) break;
#pragma unroll
    for (int j = 0; j < kSmallC; ++j) {
      if (j >= nsmall) break;
      float v = acc[ci][j];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) {
        const int o = big_is_gy ? c : j, i = big_is_gy ? j : c;
        atomicAdd(gw + (int64_t)o * cin + i, v);
      }
    }
    if (gb && big_is_gy) {
      float v = bsum[ci];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) atomicAdd(gb + c, v);
    }
  }
  if (gb && !big_is_gy && warp == 0) {
#pragma unroll
    for (int j = 0; j < kSmallC; ++j) {
      if (j >= nsmall) break;
      float v = ssum[j];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) atomicAdd(gb + j, v);
    }
  }
}


// One channel on the small side (the lifting 1 -> C and projection C -> 1 of every architecture in the reference):
// pure streaming. blockIdx.y selects a group of WC_CG big-side channels; a thread keeps one float4 of the small side in
// registers and issues the WC_CG row loads of its 4 pixels back to back (16 x 16 B in flight per thread), accumulating
// per-channel dot products (and channel sums for the bias) across all its units; one block reduction + atomic flush.
constexpr int WC_CG = 16;
template <bool BIG_IS_GY>
__global__ void __launch_bounds__(256, 2)
pw_wgrad_c1_kernel(const float* __restrict__ big, const float* __restrict__ small_, float* __restrict__ gw,
                   float* __restrict__ gb, int batch, int nbig, int64_t n) {
  pdl_enter();
  __shared__ float red[8][2 * WC_CG + 1];
  const int c0 = blockIdx.y * WC_CG;
  const int64_t chunks = n / 1024;  // the launcher guarantees n % 1024 == 0
  const int64_t units = (int64_t)batch * chunks;
  float acc[WC_CG], bs[WC_CG], ss = 0.f;
#pragma unroll
  for (int c = 0; c < WC_CG; ++c) acc[c] = bs[c] = 0.f;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = (int)(u / chunks);
    const int64_t p = (u % chunks) * 1024 + threadIdx.x * 4;
    const float4 sv = __ldg((const float4*)(small_ + (int64_t)b * n + p));
    const float* row = big + ((int64_t)b * nbig + c0) * n + p;
    float4 v[WC_CG];
#pragma unroll
    for (int c = 0; c < WC_CG; ++c) v[c] = (c0 + c < nbig) ? __ldg((const float4*)(row + (int64_t)c * n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < WC_CG; ++c) {
      acc[c] = fmaf(v[c].x, sv.x, fmaf(v[c].y, sv.y, fmaf(v[c].z, sv.z, fmaf(v[c].w, sv.w, acc[c]))));
      if (BIG_IS_GY) bs[c] += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    }
    if (!BIG_IS_GY) ss += (sv.x + sv.y) + (sv.z + sv.w);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < WC_CG; ++c) {
    float a = acc[c], s2 = bs[c];
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      if (BIG_IS_GY) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    }
    if (lane == 0) { red[warp][c] = a; red[warp][WC_CG + c] = s2; }
  }
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (lane == 0) red[warp][2 * WC_CG] = ss;
  __syncthreads();
  if (threadIdx.x < 2 * WC_CG + 1) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int c = threadIdx.x % WC_CG;
    if (threadIdx.x < WC_CG) {
      if (c0 + c < nbig) atomicAdd(gw + c0 + c, t);  // gw is [nbig] either way: (C_o,1) or (1,C_i)
    } else if (threadIdx.x < 2 * WC_CG) {
      if (BIG_IS_GY && gb && c0 + c < nbig) atomicAdd(gb + c0 + c, t);
    } else if (!BIG_IS_GY && gb && blockIdx.y == 0) {
      atomicAdd(gb, t);
    }
  }
}

}  // namespace pdeqx

using namespace pdeqx;

extern "C" int pdeqx_pointwise_fwd(int32_t batch, int32_t c_in, int32_t c_out, int64_t n_pixels, const float* x,
                                   const float* w, const float* b, const float* add, int32_t activation, float* y,
                                   void* stream) {
  PDEQX_CHECK_ARG(x && w && y, "null argument");
  PDEQX_CHECK_ARG(batch >= 0 && c_in >= 1 && c_out >= 1 && n_pixels >= 0, "bad sizes");
  PDEQX_CHECK_ARG(activation >= 0 && activation <= 2, "unknown activation %d", activation);
  return pointwise_apply(batch, c_in, c_out, n_pixels, x, w, c_in, 1, b, add, activation, y, (cudaStream_t)stream);
}

extern "C" int pdeqx_pointwise_bwd(int32_t batch, int32_t c_in, int32_t c_out, int64_t n_pixels, const float* x,
                                   const float* w, const float* gy, float* gx, float* gw, float* gb, void* stream) {
  PDEQX_CHECK_ARG(x && w && gy && gw, "null argument");
  PDEQX_CHECK_ARG(batch >= 0 && c_in >= 1 && c_out >= 1 && n_pixels >= 0, "bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  if (gx)  // gx[b,i,n] = sum_o w[o,i] gy[b,o,n]: the forward op with W' = w^T (roles of c_in/c_out swapped)
    PDEQX_TRY(pointwise_apply(batch, c_out, c_in, n_pixels, gy, w, 1, c_in, nullptr, nullptr, PDEQX_ACT_IDENTITY, gx, s));
  PDEQX_TRY(zero_small(gw, (int64_t)c_in * c_out, gb, c_out, s));
  if (batch == 0 || n_pixels == 0) return PDEQX_OK;
  const int nbig = c_out >= c_in ? c_out : c_in, nsmall = c_out >= c_in ? c_in : c_out;
  if (nsmall == 1 && n_pixels % 1024 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)gy % 16 == 0) {
    const bool big_is_gy = c_out >= c_in;
    const int64_t units = (int64_t)batch * (n_pixels / 1024);
    const int groups = (int)ceil_div(nbig, WC_CG);
    int per_group = (148 * 2 + groups - 1) / groups;  // 2 resident CTAs per SM over all channel groups
    if (per_group > units) per_group = (int)units;
    dim3 grid(per_group, groups);
    if (big_is_gy) PDEQX_CUDA(launch_pdl(pw_wgrad_c1_kernel<true>, grid, dim3(256), 0, s, gy, x, gw, gb, batch, nbig, n_pixels));
    else PDEQX_CUDA(launch_pdl(pw_wgrad_c1_kernel<false>, grid, dim3(256), 0, s, x, gy, gw, gb, batch, nbig, n_pixels));
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
  if (nsmall <= kSmallC && nbig <= 8 * WS_MAXCH) {
    const int64_t units = (int64_t)batch * ceil_div(n_pixels, WS_CHUNK);
    const int grid = (int)(units < 148 ? units : 148);
    const size_t smem = (size_t)nsmall * WS_CHUNK * sizeof(float);
    pw_wgrad_small_kernel<<<grid, 256, smem, s>>>(gy, x, gw, gb, batch, c_in, c_out, n_pixels);
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
  const int64_t chunks_per_b = ceil_div(n_pixels, WG_BK);
  const int64_t total = (int64_t)batch * chunks_per_b;
  const int ot = (int)ceil_div(c_out, 64), it = (int)ceil_div(c_in, 64);
  int gx_blocks = (int)(total < 148 * 4 ? total : 148 * 4);
  pw_wgrad_kernel<<<dim3(gx_blocks, ot, it), 256, 0, s>>>(gy, x, gw, gb, batch, c_in, c_out, n_pixels, chunks_per_b);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}
