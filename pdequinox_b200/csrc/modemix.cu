// Per-mode complex channel mix of SpectralConv and its two adjoints.
//
//   replaces the einsum "i...,oi...->o..." over the 2^(D-1) corner blocks
//   (pdequinox/conv/_spectral_conv.py:129,133-135).
//
// Mode space is stored as [B, C, J] complex-interleaved with J = (2m_0, .., 2m_{D-2}, m_{D-1})
// flattened (last axis fastest). The weights keep the reference layout
// (G, o, i, m_0, .., m_{D-1}), real and imaginary parts in separate arrays; corner g is
// selected by which half of each leading axis a retained column j falls in (bit a of g).
#include "common.cuh"
#include "spectral_plan.h"

namespace pdeqx {

struct ModeMap {
  int D;
  int m[PDEQX_MAX_DIMS];
  int64_t mprod;  // prod m_a
};

// p -> (g, loc): corner index and offset inside the corner's (m_0..m_{D-1}) block
__device__ __forceinline__ void mode_to_weight(const ModeMap& mm, int64_t p, int& g, int64_t& loc) {
  const int D = mm.D;
  int64_t k = p % mm.m[D - 1];
  int64_t rest = p / mm.m[D - 1];
  g = 0;
  loc = k;
  int64_t stride = mm.m[D - 1];
  for (int a = D - 2; a >= 0; --a) {
    int j = (int)(rest % (2 * mm.m[a]));
    rest /= 2 * mm.m[a];
    int hi = j >= mm.m[a];
    g |= hi << a;
    loc += (int64_t)(j - hi * mm.m[a]) * stride;
    stride *= mm.m[a];
  }
}

// out[b,u,p] = sum_v W(u,v,p) * in[b,v,p]
template <bool TC>
__global__ void __launch_bounds__(256)
mode_mix_kernel(const float2* __restrict__ in, const float* __restrict__ wr, const float* __restrict__ wi,
                float2* __restrict__ out, int B, int U, int V, int64_t J, int O, int I, ModeMap mm) {
  const int tid = threadIdx.x;
  const int pl = tid & 15, bg = (tid >> 4) & 3, ug = tid >> 6;
  const int64_t p = (int64_t)blockIdx.x * 16 + pl;
  const int b0 = blockIdx.y * 16 + bg * 4;
  const int u0 = blockIdx.z * 16 + ug * 4;
  if (p >= J) return;
  int g;
  int64_t loc;
  mode_to_weight(mm, p, g, loc);
  float2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  const int64_t wbase = (int64_t)g * O * I * mm.mprod + loc;
  for (int v = 0; v < V; ++v) {
    float2 x[4];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      int b = b0 + bb;
      x[bb] = (b < B) ? __ldg(in + ((int64_t)b * V + v) * J + p) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
      int u = u0 + uu;
      float a = 0.f, c = 0.f;
      if (u < U) {
        int o = TC ? v : u, i = TC ? u : v;
        int64_t off = wbase + ((int64_t)o * I + i) * mm.mprod;
        a = __ldg(wr + off);
        c = __ldg(wi + off);
        if (TC) c = -c;
      }
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        acc[bb][uu].x = fmaf(a, x[bb].x, fmaf(-c, x[bb].y, acc[bb][uu].x));
        acc[bb][uu].y = fmaf(a, x[bb].y, fmaf(c, x[bb].x, acc[bb][uu].y));
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < 4; ++bb) {
    int b = b0 + bb;
    if (b >= B) continue;
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
      int u = u0 + uu;
      if (u < U) out[((int64_t)b * U + u) * J + p] = acc[bb][uu];
    }
  }
}

// gw[g,o,i,loc] = sum_b ghat[b,o,p] * conj(xhat[b,i,p])
__global__ void __launch_bounds__(256)
mode_mix_dw_kernel(const float2* __restrict__ ghat, const float2* __restrict__ xhat, float* __restrict__ gwr,
                   float* __restrict__ gwi, int B, int64_t J, int O, int I, ModeMap mm) {
  const int tid = threadIdx.x;
  const int pl = tid & 15, og = (tid >> 4) & 3, ig = tid >> 6;
  const int64_t p = (int64_t)blockIdx.x * 16 + pl;
  const int o0 = blockIdx.y * 16 + og * 4;
  const int i0 = blockIdx.z * 16 + ig * 4;
  if (p >= J) return;
  float2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    float2 gv[4], xv[4];
#pragma unroll
    for (int oo = 0; oo < 4; ++oo)
      gv[oo] = (o0 + oo < O) ? __ldg(ghat + ((int64_t)b * O + o0 + oo) * J + p) : make_float2(0.f, 0.f);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
      xv[ii] = (i0 + ii < I) ? __ldg(xhat + ((int64_t)b * I + i0 + ii) * J + p) : make_float2(0.f, 0.f);
#pragma unroll
    for (int oo = 0; oo < 4; ++oo)
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        acc[oo][ii].x = fmaf(gv[oo].x, xv[ii].x, fmaf(gv[oo].y, xv[ii].y, acc[oo][ii].x));
        acc[oo][ii].y = fmaf(gv[oo].y, xv[ii].x, fmaf(-gv[oo].x, xv[ii].y, acc[oo][ii].y));
      }
  }
  int g;
  int64_t loc;
  mode_to_weight(mm, p, g, loc);
  const int64_t wbase = (int64_t)g * O * I * mm.mprod + loc;
#pragma unroll
  for (int oo = 0; oo < 4; ++oo) {
    if (o0 + oo >= O) continue;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      if (i0 + ii >= I) continue;
      int64_t off = wbase + ((int64_t)(o0 + oo) * I + i0 + ii) * mm.mprod;
      gwr[off] = acc[oo][ii].x;
      gwi[off] = acc[oo][ii].y;
    }
  }
}

static ModeMap make_map(const pdeqx_spectral_plan* p) {
  ModeMap mm;
  mm.D = p->desc.ndim;
  mm.mprod = 1;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    mm.m[a] = a < mm.D ? p->desc.modes[a] : 1;
    if (a < mm.D) mm.mprod *= p->desc.modes[a];
  }
  return mm;
}

int mode_mix(const pdeqx_spectral_plan* p, const float* in, const float* w_re, const float* w_im,
             bool transposed_conj, float* out, cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  const int O = d.c_out, I = d.c_in;
  const int U = transposed_conj ? I : O, V = transposed_conj ? O : I;
  dim3 grid((unsigned)ceil_div(p->J, 16), (unsigned)ceil_div(d.batch, 16), (unsigned)ceil_div(U, 16));
  PDEQX_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "mode_mix: grid too large");
  ModeMap mm = make_map(p);
  if (transposed_conj)
    mode_mix_kernel<true><<<grid, 256, 0, s>>>((const float2*)in, w_re, w_im, (float2*)out, d.batch, U, V, p->J, O, I, mm);
  else
    mode_mix_kernel<false><<<grid, 256, 0, s>>>((const float2*)in, w_re, w_im, (float2*)out, d.batch, U, V, p->J, O, I, mm);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

int mode_mix_dw(const pdeqx_spectral_plan* p, const float* ghat, const float* xhat, float* gw_re, float* gw_im,
                cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  dim3 grid((unsigned)ceil_div(p->J, 16), (unsigned)ceil_div(d.c_out, 16), (unsigned)ceil_div(d.c_in, 16));
  PDEQX_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "mode_mix_dw: grid too large");
  mode_mix_dw_kernel<<<grid, 256, 0, s>>>((const float2*)ghat, (const float2*)xhat, gw_re, gw_im, d.batch, p->J,
                                           d.c_out, d.c_in, make_map(p));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx
