// GroupNorm as used by DilatedResBlock / ClassicDoubleConvBlock / ClassicResBlock(use_norm=True).
//
//   replaces eqx.nn.GroupNorm(groups, channels, eps=1e-5) (third-party; call sites
//   pdequinox/blocks/_dilated_res_block.py:88-99, _classic_double_conv_block.py:75-83,
//   _classic_res_block.py:80-89): per sample and group, mean and BIASED variance over
//   (channels/groups x spatial); y = (x - mean) * rsqrt(var + eps) * weight[c] + bias[c].
//
// HBM-bound streaming kernels: statistics are accumulated per (sample, channel) in one pass
// (fp32 per thread, fp64 across threads), the normalisation pass re-reads x (L2-resident for the
// hot-path sizes is not guaranteed: 2 reads + 1 write per element is the algorithmic traffic).
#include "common.cuh"

namespace pdeqx {

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// partial[(b*C + c)*2 + {0,1}] = sum x, sum x^2 over the channel's n points. grid (B*C), 256 threads.
__global__ void __launch_bounds__(256)
gn_channel_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, int64_t n) {
  __shared__ double red[8];
  const float* xc = x + (int64_t)blockIdx.x * n;
  float s = 0.f, q = 0.f;
  double ds = 0.0, dq = 0.0;
  int cnt = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float v = __ldg(xc + i);
    s += v;
    q = fmaf(v, v, q);
    if (++cnt == 256) {  // bound the fp32 accumulation length
      ds += s; dq += q; s = 0.f; q = 0.f; cnt = 0;
    }
  }
  ds += s;
  dq += q;
  double ts = block_sum_d(ds, red), tq = block_sum_d(dq, red);
  if (threadIdx.x == 0) {
    partial[(int64_t)blockIdx.x * 2] = ts;
    partial[(int64_t)blockIdx.x * 2 + 1] = tq;
  }
}

// stats[(b*G + g)*2 + {0,1}] = mean, rstd; one warp per (sample, group), lanes split the group's channels
__global__ void gn_finalize_kernel(const double* __restrict__ partial, float* __restrict__ stats, int BG, int cpg,
                                   int64_t n, float eps) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= BG) return;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < cpg; c += 32) {
    s += partial[((int64_t)i * cpg + c) * 2];
    q += partial[((int64_t)i * cpg + c) * 2 + 1];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if (lane != 0) return;
  const double cnt = (double)cpg * (double)n;
  const double mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// grid (B*C, chunks)
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ y, int C, int cpg, int G, int64_t n) {
  const int bc = blockIdx.x, b = bc / C, c = bc % C, g = c / cpg;
  const float mean = stats[2 * (b * G + g)], rstd = stats[2 * (b * G + g) + 1];
  const float sc = rstd * (w ? __ldg(w + c) : 1.f);
  const float sh = (bias ? __ldg(bias + c) : 0.f) - mean * sc;
  const float* xc = x + (int64_t)bc * n;
  float* yc = y + (int64_t)bc * n;
  if ((n & 3) == 0) {
    const int64_t nv = n / 4;
    for (int64_t v = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.y * blockDim.x) {
      float4 t = __ldg((const float4*)xc + v);
      ((float4*)yc)[v] = make_float4(fmaf(t.x, sc, sh), fmaf(t.y, sc, sh), fmaf(t.z, sc, sh), fmaf(t.w, sc, sh));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x)
      yc[i] = fmaf(__ldg(xc + i), sc, sh);
  }
}

// per (b,c): scratch[(b*C+c)*2 + {0,1}] = sum gy, sum gy*xhat
__global__ void __launch_bounds__(256)
gn_bwd_channel_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ stats,
                      float* __restrict__ scratch, int C, int cpg, int G, int64_t n) {
  __shared__ double red[8];
  const int bc = blockIdx.x, b = bc / C, c = bc % C, g = c / cpg;
  const float mean = stats[2 * (b * G + g)], rstd = stats[2 * (b * G + g) + 1];
  const float* xc = x + (int64_t)bc * n;
  const float* gc = gy + (int64_t)bc * n;
  float s = 0.f, q = 0.f;
  double ds = 0.0, dq = 0.0;
  int cnt = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float gv = __ldg(gc + i), xh = (__ldg(xc + i) - mean) * rstd;
    s += gv;
    q = fmaf(gv, xh, q);
    if (++cnt == 256) {
      ds += s; dq += q; s = 0.f; q = 0.f; cnt = 0;
    }
  }
  ds += s;
  dq += q;
  double ts = block_sum_d(ds, red), tq = block_sum_d(dq, red);
  if (threadIdx.x == 0) {
    scratch[(int64_t)bc * 2] = (float)ts;
    scratch[(int64_t)bc * 2 + 1] = (float)tq;
  }
}

// gx = rstd * (w gy - s1 - xhat s2), s1 = mean_group(w gy), s2 = mean_group(w gy xhat). grid (B*C, chunks)
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ stats,
                    const float* __restrict__ w, const float* __restrict__ scratch, const float* __restrict__ mask,
                    float* __restrict__ gx, int C, int cpg, int G, int64_t n) {
  const int bc = blockIdx.x, b = bc / C, c = bc % C, g = c / cpg;
  const float mean = stats[2 * (b * G + g)], rstd = stats[2 * (b * G + g) + 1];
  // group sums of the per-channel partials: the block's threads split the group's channels (a serial loop per thread over
  // up to 256 channels made this kernel 56 us on 131 KB tensors), then a two-level shuffle / shared-memory reduction
  __shared__ float red1[8], red2[8];
  float s1 = 0.f, s2 = 0.f;
  for (int cc = threadIdx.x; cc < cpg; cc += blockDim.x) {
    const int ch = g * cpg + cc;
    const float wv = w ? __ldg(w + ch) : 1.f;
    s1 = fmaf(wv, scratch[((int64_t)b * C + ch) * 2], s1);
    s2 = fmaf(wv, scratch[((int64_t)b * C + ch) * 2 + 1], s2);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if ((threadIdx.x & 31) == 0) {
    red1[threadIdx.x >> 5] = s1;
    red2[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  s1 = s2 = 0.f;
#pragma unroll
  for (int wv2 = 0; wv2 < 8; ++wv2) {
    s1 += red1[wv2];
    s2 += red2[wv2];
  }
  const float inv = 1.f / ((float)cpg * (float)n);
  s1 *= inv;
  s2 *= inv;
  const float wv = w ? __ldg(w + c) : 1.f;
  const float* xc = x + (int64_t)bc * n;
  const float* gc = gy + (int64_t)bc * n;
  float* oc = gx + (int64_t)bc * n;
  const float* mc = mask ? mask + (int64_t)bc * n : nullptr;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x) {
    float xh = (__ldg(xc + i) - mean) * rstd;
    float v = rstd * (wv * __ldg(gc + i) - s1 - xh * s2);
    if (mc && __ldg(mc + i) <= 0.f) v = 0.f;  // relu' of the layer that produced x (its output IS x when mask == x)
    oc[i] = v;
  }
}

// gw[c] = sum_b scratch[b,c,1]; gb[c] = sum_b scratch[b,c,0]
__global__ void gn_bwd_param_kernel(const float* __restrict__ scratch, float* __restrict__ gw, float* __restrict__ gb,
                                    int B, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, d = 0.f;
  for (int b = 0; b < B; ++b) {
    a += scratch[((int64_t)b * C + c) * 2];
    d += scratch[((int64_t)b * C + c) * 2 + 1];
  }
  if (gb) gb[c] = a;
  if (gw) gw[c] = d;
}

}  // namespace pdeqx

using namespace pdeqx;

static int chunks_for(int64_t n, int bc) {
  int64_t want = ceil_div((int64_t)148 * 8, bc > 0 ? bc : 1);
  int64_t maxc = ceil_div(n, 1024);
  int64_t c = want < maxc ? want : maxc;
  return (int)(c < 1 ? 1 : (c > 65535 ? 65535 : c));
}

extern "C" size_t pdeqx_groupnorm_workspace_bytes(int32_t batch, int32_t channels) {
  return (size_t)batch * channels * 2 * sizeof(double);
}

extern "C" int pdeqx_groupnorm_fwd(int32_t batch, int32_t channels, int32_t groups, int64_t n, const float* x,
                                   const float* weight, const float* bias, float eps, float* y, float* stats,
                                   void* workspace, void* stream) {
  PDEQX_CHECK_ARG(x && y && stats && workspace, "null argument");
  PDEQX_CHECK_ARG(batch >= 0 && channels >= 1 && groups >= 1 && n >= 1, "bad sizes");
  PDEQX_CHECK_ARG(channels % groups == 0, "The number of groups must divide the number of channels.");
  PDEQX_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)workspace & 7) == 0, "misaligned pointer");
  if (batch == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int BC = batch * channels, BG = batch * groups, cpg = channels / groups;
  double* partial = (double*)workspace;
  gn_channel_stats_kernel<<<BC, 256, 0, s>>>(x, partial, n);
  PDEQX_LAUNCHED();
  gn_finalize_kernel<<<(int)ceil_div(BG, 4), 128, 0, s>>>(partial, stats, BG, cpg, n, eps);
  PDEQX_LAUNCHED();
  gn_apply_kernel<<<dim3(BC, chunks_for(n, BC)), 256, 0, s>>>(x, stats, weight, bias, y, channels, cpg, groups, n);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_groupnorm_bwd(int32_t batch, int32_t channels, int32_t groups, int64_t n, const float* x,
                                   const float* weight, const float* stats, const float* gy, const float* relu_mask,
                                   float* gx, float* gweight, float* gbias, float* scratch, void* stream) {
  PDEQX_CHECK_ARG(x && stats && gy && gx && scratch, "null argument");
  PDEQX_CHECK_ARG(batch >= 0 && channels >= 1 && groups >= 1 && n >= 1 && channels % groups == 0, "bad sizes");
  if (batch == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int BC = batch * channels, cpg = channels / groups;
  gn_bwd_channel_kernel<<<BC, 256, 0, s>>>(x, gy, stats, scratch, channels, cpg, groups, n);
  PDEQX_LAUNCHED();
  gn_bwd_apply_kernel<<<dim3(BC, chunks_for(n, BC)), 256, 0, s>>>(x, gy, stats, weight, scratch, relu_mask, gx, channels, cpg, groups, n);
  PDEQX_LAUNCHED();
  if (gweight || gbias) {
    gn_bwd_param_kernel<<<(int)ceil_div(channels, 128), 128, 0, s>>>(scratch, gweight, gbias, batch, channels);
    PDEQX_LAUNCHED();
  }
  return PDEQX_OK;
}
