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

__device__ __forceinline__ float to_tf32_nearest(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

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
// 128-bit loads, four independent load / accumulate chains per thread (one scalar load per loop trip ran at a third of
// the HBM bandwidth); fp32 per chain (<= n / 1024 terms of 4), fp64 across chains and threads.
__global__ void __launch_bounds__(256)
gn_channel_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, int64_t n) {
  __shared__ double red[8];
  const float* xc = x + (int64_t)blockIdx.x * n;
  double ds = 0.0, dq = 0.0;
  if ((n & 3) == 0 && (((uintptr_t)xc) & 15) == 0) {
    const float4* xv = (const float4*)xc;
    const int64_t nv = n / 4;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    int64_t i = threadIdx.x;
    for (; i + 3 * 256 < nv; i += 4 * 256) {
      float4 t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = __ldg(xv + i + j * 256);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] += (t[j].x + t[j].y) + (t[j].z + t[j].w);
        q[j] = fmaf(t[j].x, t[j].x, fmaf(t[j].y, t[j].y, fmaf(t[j].z, t[j].z, fmaf(t[j].w, t[j].w, q[j]))));
      }
      if (++cnt == 64) {  // bound the fp32 accumulation length
#pragma unroll
        for (int j = 0; j < 4; ++j) { ds += s[j]; dq += q[j]; s[j] = 0.f; q[j] = 0.f; }
        cnt = 0;
      }
    }
    for (; i < nv; i += 256) {
      const float4 t = __ldg(xv + i);
      s[0] += (t.x + t.y) + (t.z + t.w);
      q[0] = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, q[0]))));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { ds += s[j]; dq += q[j]; }
  } else {
    float s = 0.f, q = 0.f;
    int cnt = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      float v = __ldg(xc + i);
      s += v;
      q = fmaf(v, v, q);
      if (++cnt == 256) {
        ds += s; dq += q; s = 0.f; q = 0.f; cnt = 0;
      }
    }
    ds += s;
    dq += q;
  }
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

// per (b,c): scratch[(b*C+c)*2 + {0,1}] = sum gy, sum gy*xhat   (128-bit loads, two independent chains per thread)
__global__ void __launch_bounds__(256)
gn_bwd_channel_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ stats,
                      float* __restrict__ scratch, int C, int cpg, int G, int64_t n) {
  __shared__ double red[8];
  const int bc = blockIdx.x, b = bc / C, c = bc % C, g = c / cpg;
  const float mean = stats[2 * (b * G + g)], rstd = stats[2 * (b * G + g) + 1];
  const float* xc = x + (int64_t)bc * n;
  const float* gc = gy + (int64_t)bc * n;
  double ds = 0.0, dq = 0.0;
  if ((n & 3) == 0 && ((((uintptr_t)xc) | ((uintptr_t)gc)) & 15) == 0) {
    const float4* xv = (const float4*)xc;
    const float4* gv = (const float4*)gc;
    const int64_t nv = n / 4;
    float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
    int cnt = 0;
    int64_t i = threadIdx.x;
    for (; i + 256 < nv; i += 2 * 256) {
      float4 a[2], t[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) { a[j] = __ldg(gv + i + j * 256); t[j] = __ldg(xv + i + j * 256); }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        s[j] += (a[j].x + a[j].y) + (a[j].z + a[j].w);
        q[j] = fmaf(a[j].x, (t[j].x - mean) * rstd, fmaf(a[j].y, (t[j].y - mean) * rstd,
               fmaf(a[j].z, (t[j].z - mean) * rstd, fmaf(a[j].w, (t[j].w - mean) * rstd, q[j]))));
      }
      if (++cnt == 64) {
        ds += (double)s[0] + (double)s[1]; dq += (double)q[0] + (double)q[1];
        s[0] = s[1] = q[0] = q[1] = 0.f; cnt = 0;
      }
    }
    for (; i < nv; i += 256) {
      const float4 a = __ldg(gv + i), t = __ldg(xv + i);
      s[0] += (a.x + a.y) + (a.z + a.w);
      q[0] = fmaf(a.x, (t.x - mean) * rstd, fmaf(a.y, (t.y - mean) * rstd, fmaf(a.z, (t.z - mean) * rstd, fmaf(a.w, (t.w - mean) * rstd, q[0]))));
    }
    ds += (double)s[0] + (double)s[1];
    dq += (double)q[0] + (double)q[1];
  } else {
    float s = 0.f, q = 0.f;
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
  }
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

// ------------------------------------------------------------------------------------------
// Small groups (cpg * n <= 4096 elements: every level of ClassicUNet 1D, configs[1]): ONE block per (sample, group) keeps the
// group in registers -- statistics (two-pass: mean, then centred squares), normalisation and the saved (mean, rstd) in one
// launch instead of three; the reverse pass in one launch (+ the parameter reduction over the batch) instead of three. The
// per-(b, c) kernels above launch B*C blocks of 256 threads for n = 4..64 points and were 30 % of the UNet train step.
// ------------------------------------------------------------------------------------------
constexpr int kGnSmallMax = 4096;
constexpr int kGnSmallPer = kGnSmallMax / 256;

__device__ __forceinline__ float block_sum_f(float v, float* red) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256)
gn_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                    float* __restrict__ stats, int C, int cpg, int G, int n, float eps) {
  __shared__ float red[8];
  const int bg = blockIdx.x, b = bg / G, g = bg % G, m = cpg * n;
  const float* xg = x + ((int64_t)b * C + (int64_t)g * cpg) * n;
  float* yg = y + ((int64_t)b * C + (int64_t)g * cpg) * n;
  float v[kGnSmallPer];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kGnSmallPer; ++k) {
    const int e = threadIdx.x + k * 256;
    v[k] = e < m ? __ldg(xg + e) : 0.f;
    s += v[k];
  }
  const float mean = block_sum_f(s, red) / (float)m;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kGnSmallPer; ++k) {
    const int e = threadIdx.x + k * 256;
    const float d = e < m ? v[k] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(block_sum_f(q, red) / (float)m + eps);
  if (threadIdx.x == 0) {
    stats[2 * bg] = mean;
    stats[2 * bg + 1] = rstd;
  }
#pragma unroll
  for (int k = 0; k < kGnSmallPer; ++k) {
    const int e = threadIdx.x + k * 256;
    if (e < m) {
      const int c = g * cpg + e / n;
      yg[e] = fmaf((v[k] - mean) * rstd, w ? __ldg(w + c) : 1.f, bias ? __ldg(bias + c) : 0.f);
    }
  }
}

// gx = rstd (w gy - s1 - xhat s2) [* (mask > 0)]; scratch[(b*C + c)*2 + {0,1}] = sum gy, sum gy xhat (deterministic: every
// product is parked in shared memory and thread c adds up its channel's n entries in order)
__global__ void __launch_bounds__(256)
gn_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ stats,
                    const float* __restrict__ w, const float* __restrict__ mask, float* __restrict__ gx, float* __restrict__ scratch,
                    int C, int cpg, int G, int n) {
  __shared__ float red[8];
  __shared__ float pg[kGnSmallMax], px[kGnSmallMax];
  const int bg = blockIdx.x, b = bg / G, g = bg % G, m = cpg * n;
  const int64_t base = ((int64_t)b * C + (int64_t)g * cpg) * n;
  const float mean = stats[2 * bg], rstd = stats[2 * bg + 1];
  float xh[kGnSmallPer], wg[kGnSmallPer];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < kGnSmallPer; ++k) {
    const int e = threadIdx.x + k * 256;
    xh[k] = 0.f;
    wg[k] = 0.f;
    if (e < m) {
      const float gv = __ldg(gy + base + e);
      xh[k] = (__ldg(x + base + e) - mean) * rstd;
      pg[e] = gv;
      px[e] = gv * xh[k];
      wg[k] = gv * (w ? __ldg(w + g * cpg + e / n) : 1.f);
      s1 += wg[k];
      s2 = fmaf(wg[k], xh[k], s2);
    }
  }
  s1 = block_sum_f(s1, red) / (float)m;
  s2 = block_sum_f(s2, red) / (float)m;  // (the barriers inside also publish pg / px)
#pragma unroll
  for (int k = 0; k < kGnSmallPer; ++k) {
    const int e = threadIdx.x + k * 256;
    if (e < m) {
      float r = rstd * (wg[k] - s1 - xh[k] * s2);
      if (mask && __ldg(mask + base + e) <= 0.f) r = 0.f;
      gx[base + e] = r;
    }
  }
  for (int c = threadIdx.x; c < cpg; c += 256) {
    float a = 0.f, d = 0.f;
    for (int i = 0; i < n; ++i) {
      a += pg[c * n + i];
      d += px[c * n + i];
    }
    scratch[((int64_t)b * C + g * cpg + c) * 2] = a;
    scratch[((int64_t)b * C + g * cpg + c) * 2 + 1] = d;
  }
}

// ------------------------------------------------------------------------------------------
// groups = 1 GroupNorm folded into the convolutions around it (tc_conv.cu): what is left for this file
// ------------------------------------------------------------------------------------------
// stats[b] = (mean, rstd) from the per-sample partial sums (sum x, sum x^2) that the PRODUCER of x wrote in its epilogue
// (parts [B][slots][2]); one warp per sample, fp64 across the slots.
__global__ void gn_parts_finalize_kernel(const float* __restrict__ parts, float* __restrict__ stats, int B, int slots,
                                         double count, float eps) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < slots; i += 32) {
    s += (double)parts[((int64_t)b * slots + i) * 2];
    q += (double)parts[((int64_t)b * slots + i) * 2 + 1];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if (lane != 0) return;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * b] = (float)mean;
  stats[2 * b + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// out = a + b with the per-sample partial sums (sum, sum of squares) of the result: grid (slots, B), the block takes one
// 1/slots share of the sample. The block output of a residual block feeds the first norm of the next one.
__global__ void __launch_bounds__(256)
add_parts_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, float* __restrict__ parts,
                 int64_t per_sample, int64_t share) {
  __shared__ double red[8];
  const int64_t base = (int64_t)blockIdx.y * per_sample + (int64_t)blockIdx.x * share;
  const float4* av = (const float4*)(a + base);
  const float4* bv = (const float4*)(b + base);
  float4* ov = (float4*)(out + base);
  float s = 0.f, q = 0.f;
  for (int64_t i = threadIdx.x; i < share / 4; i += 256) {
    const float4 x = __ldg(av + i), y = __ldg(bv + i);
    // (rounded to tf32-nearest like every other tensor a tcgen05 stage reads next: the next block's first convolution and
    // its filter gradient would otherwise truncate it, after the statistics below were taken from the untruncated values)
    const float4 r = make_float4(to_tf32_nearest(x.x + y.x), to_tf32_nearest(x.y + y.y), to_tf32_nearest(x.z + y.z),
                                 to_tf32_nearest(x.w + y.w));
    ov[i] = r;
    s += (r.x + r.y) + (r.z + r.w);
    q = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, fmaf(r.w, r.w, q))));
  }
  const double ts = block_sum_d((double)s, red), tq = block_sum_d((double)q, red);
  if (threadIdx.x == 0) {
    float* dst = parts + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    dst[0] = (float)ts;
    dst[1] = (float)tq;
  }
}

// Reverse pass of the folded norm (groups = 1). v = cotangent of the normalised tensor (written by the data-gradient kernel,
// whose epilogue also summed P1 = sum_c gamma_c sum_p v and P2 = sum_c gamma_c sum_p v h per sample into `parts`):
//   gx = rstd (gamma_c v - s1 - xhat s2) [* relu'(h)] [+ add],   s1 = P1 / (C n),  s2 = rstd (P2 - mu P1) / (C n)
// and, in the same pass over v and h, the per-(b, c) sums (sum v, sum v xhat) that give the norm's parameter gradients.
// grid (B*C), 256 threads; n % 4 == 0.
__global__ void __launch_bounds__(256)
gn_bwd_fused_kernel(const float* __restrict__ h, const float* __restrict__ v, const float* __restrict__ stats,
                    const float* __restrict__ w, const float* __restrict__ parts, int slots, int mask_h,
                    const float* __restrict__ add, float* __restrict__ gx, float* __restrict__ scratch, int C, int64_t n) {
  __shared__ double red[8];
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  double p1 = 0.0, p2 = 0.0;
  for (int i = threadIdx.x; i < slots; i += 256) {
    p1 += (double)parts[((int64_t)b * slots + i) * 2];
    p2 += (double)parts[((int64_t)b * slots + i) * 2 + 1];
  }
  p1 = block_sum_d(p1, red);
  p2 = block_sum_d(p2, red);
  const double inv = 1.0 / ((double)C * (double)n);
  const float s1 = (float)(p1 * inv), s2 = (float)((double)rstd * (p2 - (double)mean * p1) * inv);
  const float wv = w ? __ldg(w + c) : 1.f;
  const float4* hv = (const float4*)(h + (int64_t)bc * n);
  const float4* vv = (const float4*)(v + (int64_t)bc * n);
  const float4* av = add ? (const float4*)(add + (int64_t)bc * n) : nullptr;
  float4* ov = (float4*)(gx + (int64_t)bc * n);
  float sv[2] = {0.f, 0.f}, sx[2] = {0.f, 0.f};
  const int64_t nv = n / 4;
  auto one = [&](const float4 hh, const float4 gg, const float4 aa, float& acc_v, float& acc_x) {
    float hi[4] = {hh.x, hh.y, hh.z, hh.w}, gi[4] = {gg.x, gg.y, gg.z, gg.w}, ai[4] = {aa.x, aa.y, aa.z, aa.w}, o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (hi[e] - mean) * rstd;
      float r = rstd * (fmaf(wv, gi[e], -s1) - xh * s2);
      if (mask_h && hi[e] <= 0.f) r = 0.f;  // relu' of the layer whose output h is
      // the consumers are tcgen05 stages (filter / data gradient of the layer in front), which would TRUNCATE fp32 to tf32:
      // round to nearest here, like every other producer of a tensor-core operand
      o[e] = to_tf32_nearest(r + ai[e]);
      acc_v += gi[e];
      acc_x = fmaf(gi[e], xh, acc_x);
    }
    return make_float4(o[0], o[1], o[2], o[3]);
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t i = threadIdx.x;
  for (; i + 256 < nv; i += 2 * 256) {
    const float4 h0 = __ldg(hv + i), h1 = __ldg(hv + i + 256), g0 = __ldg(vv + i), g1 = __ldg(vv + i + 256);
    const float4 a0 = av ? __ldg(av + i) : zero4, a1 = av ? __ldg(av + i + 256) : zero4;
    ov[i] = one(h0, g0, a0, sv[0], sx[0]);
    ov[i + 256] = one(h1, g1, a1, sv[1], sx[1]);
  }
  for (; i < nv; i += 256) ov[i] = one(__ldg(hv + i), __ldg(vv + i), av ? __ldg(av + i) : zero4, sv[0], sx[0]);
  const double tv = block_sum_d((double)sv[0] + (double)sv[1], red), tx = block_sum_d((double)sx[0] + (double)sx[1], red);
  if (threadIdx.x == 0) {
    scratch[(int64_t)bc * 2] = (float)tv;
    scratch[(int64_t)bc * 2 + 1] = (float)tx;
  }
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
  if ((int64_t)cpg * n <= kGnSmallMax) {  // the whole group in one block's registers: one launch
    gn_small_fwd_kernel<<<BG, 256, 0, s>>>(x, weight, bias, y, stats, channels, cpg, groups, (int)n, eps);
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
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
  if ((int64_t)cpg * n <= kGnSmallMax) {
    gn_small_bwd_kernel<<<batch * groups, 256, 0, s>>>(x, gy, stats, weight, relu_mask, gx, scratch, channels, cpg, groups, (int)n);
    PDEQX_LAUNCHED();
  } else {
    gn_bwd_channel_kernel<<<BC, 256, 0, s>>>(x, gy, stats, scratch, channels, cpg, groups, n);
    PDEQX_LAUNCHED();
    gn_bwd_apply_kernel<<<dim3(BC, chunks_for(n, BC)), 256, 0, s>>>(x, gy, stats, weight, scratch, relu_mask, gx, channels, cpg, groups, n);
    PDEQX_LAUNCHED();
  }
  if (gweight || gbias) {
    gn_bwd_param_kernel<<<(int)ceil_div(channels, 128), 128, 0, s>>>(scratch, gweight, gbias, batch, channels);
    PDEQX_LAUNCHED();
  }
  return PDEQX_OK;
}

/* statistics only (any number of groups): stats [B,groups,2] = (mean, rstd) */
extern "C" int pdeqx_groupnorm_stats(int32_t batch, int32_t channels, int32_t groups, int64_t n, const float* x, float eps,
                                     float* stats, void* workspace, void* stream) {
  PDEQX_CHECK_ARG(x && stats && workspace, "null argument");
  PDEQX_CHECK_ARG(batch >= 0 && channels >= 1 && groups >= 1 && n >= 1, "bad sizes");
  PDEQX_CHECK_ARG(channels % groups == 0, "The number of groups must divide the number of channels.");
  if (batch == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int BC = batch * channels, BG = batch * groups, cpg = channels / groups;
  double* partial = (double*)workspace;
  gn_channel_stats_kernel<<<BC, 256, 0, s>>>(x, partial, n);
  PDEQX_LAUNCHED();
  gn_finalize_kernel<<<(int)ceil_div(BG, 4), 128, 0, s>>>(partial, stats, BG, cpg, n, eps);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_groupnorm_stats_from_parts(int32_t batch, int32_t slots, int64_t count, const float* parts, float eps,
                                                float* stats, void* stream) {
  PDEQX_CHECK_ARG(parts && stats && batch >= 0 && slots >= 1 && count >= 1, "bad arguments");
  if (batch == 0) return PDEQX_OK;
  gn_parts_finalize_kernel<<<(int)ceil_div(batch, 4), 128, 0, (cudaStream_t)stream>>>(parts, stats, batch, slots, (double)count, eps);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_add_parts(int32_t batch, int64_t per_sample, int32_t slots, const float* a, const float* b, float* out,
                               float* parts, void* stream) {
  PDEQX_CHECK_ARG(a && b && out && parts && batch >= 0 && slots >= 1 && slots <= 65535, "bad arguments");
  PDEQX_CHECK_ARG(per_sample % ((int64_t)slots * 4) == 0, "per_sample must be a multiple of 4 * slots");
  PDEQX_CHECK_ARG(((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0, "misaligned pointer");
  if (batch == 0) return PDEQX_OK;
  add_parts_kernel<<<dim3(slots, batch), 256, 0, (cudaStream_t)stream>>>(a, b, out, parts, per_sample, per_sample / slots);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_groupnorm_bwd_fused(int32_t batch, int32_t channels, int64_t n, const float* h, const float* v,
                                         const float* weight, const float* stats, const float* parts, int32_t slots,
                                         int32_t relu_mask_h, const float* add, float* gx, float* gweight, float* gbias,
                                         float* scratch, void* stream) {
  PDEQX_CHECK_ARG(h && v && stats && parts && gx && scratch, "null argument");
  PDEQX_CHECK_ARG(batch >= 0 && channels >= 1 && n >= 4 && (n & 3) == 0 && slots >= 1, "bad sizes");
  PDEQX_CHECK_ARG(((((uintptr_t)h) | ((uintptr_t)v) | ((uintptr_t)gx) | ((uintptr_t)add)) & 15) == 0, "misaligned pointer");
  if (batch == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  gn_bwd_fused_kernel<<<batch * channels, 256, 0, s>>>(h, v, stats, weight, parts, slots, relu_mask_h, add, gx, scratch, channels, n);
  PDEQX_LAUNCHED();
  if (gweight || gbias) {
    gn_bwd_param_kernel<<<(int)ceil_div(channels, 128), 128, 0, s>>>(scratch, gweight, gbias, batch, channels);
    PDEQX_LAUNCHED();
  }
  return PDEQX_OK;
}
