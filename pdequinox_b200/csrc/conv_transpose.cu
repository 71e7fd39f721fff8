// PhysicsConvTranspose / MorePaddingConvTranspose with the reference's own semantics.
//
//   replaces pdequinox/conv/_conv.py:378-471: the weight (C_out, C_in/groups, *k) is used AS IS (no spatial flip,
//   no in/out swap) in a stride-1 cross-correlation over the input after
//     1. boundary padding by `pre` points per side (wrap / reflect / edge; zeros mode has pre = 0)  (_conv.py:420-442)
//     2. dilation by `stride` (stride-1 zeros between neighbouring points, lax lhs_dilation)           (_conv.py:459-467)
//     3. zero padding / cropping by `post` (may be negative)                                           (_conv.py:428-436,448-455)
//   None of the three intermediate arrays exists here: every tap maps its position in the virtual
//   (padded, dilated, cropped) array back to a source index, or to "structural zero".
//
// Callers: LinearConvUpBlock of ClassicUNet (k = 3, stride 2, output_padding 1; _linear_conv_up_block.py:24-36).
// Any-shape fp32 CUDA-core kernels (1-3 D, groups, dilation); the UNet up-sampling layers are <= 3 % of its work.
#include "common.cuh"

namespace pdeqx {

struct ConvTGeom {
  int ndim, batch, c_in, c_out, groups, boundary, taps;
  int spatial[PDEQX_MAX_DIMS], out[PDEQX_MAX_DIMS], vlen[PDEQX_MAX_DIMS];  // vlen: length of the padded, dilated array
  int kernel[PDEQX_MAX_DIMS], up[PDEQX_MAX_DIMS], dilation[PDEQX_MAX_DIMS];
  int pre_lo[PDEQX_MAX_DIMS], pre_hi[PDEQX_MAX_DIMS], post_lo[PDEQX_MAX_DIMS];
  int in_stride[PDEQX_MAX_DIMS], out_stride[PDEQX_MAX_DIMS];
  int64_t n_in, n_out;
};

__device__ __forceinline__ int ct_boundary_map(int q, int s, int mode) {
  if (q >= 0 && q < s) return q;
  switch (mode) {
    case PDEQX_PAD_WRAP: {
      int r = q % s;
      return r < 0 ? r + s : r;
    }
    case PDEQX_PAD_REFLECT: {
      if (s == 1) return 0;
      int period = 2 * (s - 1);
      int r = q % period;
      if (r < 0) r += period;
      return r < s ? r : period - r;
    }
    case PDEQX_PAD_EDGE:
      return q < 0 ? 0 : s - 1;
    default:
      return -1;
  }
}

// source index of output position `po`, tap `ta` on axis a, or -1 (structural zero)
__device__ __forceinline__ int ct_source(const ConvTGeom& g, int a, int po, int ta) {
  const int u = po + ta * g.dilation[a] - g.post_lo[a];  // position in the padded, dilated array
  if (u < 0 || u >= g.vlen[a] || u % g.up[a] != 0) return -1;
  return ct_boundary_map(u / g.up[a] - g.pre_lo[a], g.spatial[a], g.boundary);
}

constexpr int kCtOPT = 4;

// WARP: one warp per output quad, lanes split the input channels, shuffle reduction (the deep-channel / few-points levels of
// ClassicUNet 1D: one thread per output would loop serially over c_in x taps dependent loads -- 70 us on KB-sized tensors)
template <bool WARP>
__global__ void __launch_bounds__(256)
convT_fwd_kernel(ConvTGeom g, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y) {
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  const int ochunks = (cog + kCtOPT - 1) / kCtOPT;
  const int64_t total = (int64_t)g.batch * g.groups * ochunks * g.n_out;
  const int lane = WARP ? (int)(threadIdx.x & 31) : 0, lanes = WARP ? 32 : 1;
  const int64_t idx0 = WARP ? (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx_step = WARP ? (((int64_t)gridDim.x * blockDim.x) >> 5) : (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = idx0; idx < total; idx += idx_step) {
    const int64_t p = idx % g.n_out;
    int64_t r = idx / g.n_out;
    const int oc = (int)(r % ochunks);
    r /= ochunks;
    const int grp = (int)(r % g.groups);
    const int b = (int)(r / g.groups);
    int po[PDEQX_MAX_DIMS];
    {
      int64_t pp = p;
      for (int a = g.ndim - 1; a >= 0; --a) {
        po[a] = (int)(pp % g.out[a]);
        pp /= g.out[a];
      }
    }
    const int o0 = grp * cog + oc * kCtOPT;
    const int no = min(kCtOPT, cog - oc * kCtOPT);
    float acc[kCtOPT];
#pragma unroll
    for (int j = 0; j < kCtOPT; ++j) acc[j] = 0.f;
    const float* xb = x + ((int64_t)b * g.c_in + (int64_t)grp * cig) * g.n_in;
    for (int t = 0; t < g.taps; ++t) {
      int tt = t;
      int64_t src = 0;
      bool valid = true;
      for (int a = g.ndim - 1; a >= 0; --a) {
        const int ta = tt % g.kernel[a];
        tt /= g.kernel[a];
        const int m = ct_source(g, a, po[a], ta);
        valid = valid && (m >= 0);
        src += (int64_t)m * g.in_stride[a];
      }
      if (!valid) continue;
#pragma unroll 8
      for (int i = lane; i < cig; i += lanes) {  // independent loads in flight: the small-tensor regime is latency bound
        const float xv = __ldg(xb + (int64_t)i * g.n_in + src);
#pragma unroll
        for (int j = 0; j < kCtOPT; ++j)
          if (j < no) acc[j] = fmaf(__ldg(w + ((int64_t)(o0 + j) * cig + i) * g.taps + t), xv, acc[j]);
      }
    }
    if (WARP) {
#pragma unroll
      for (int j = 0; j < kCtOPT; ++j)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
      if (lane != 0) continue;
    }
#pragma unroll
    for (int j = 0; j < kCtOPT; ++j) {
      if (j >= no) break;
      y[((int64_t)b * g.c_out + o0 + j) * g.n_out + p] = acc[j] + (bias ? __ldg(bias + o0 + j) : 0.f);
    }
  }
}

// data gradient: gather. Input point n is read through every padded position j in [-pre_lo, s + pre_hi) that the
// boundary map sends to n (the adjoint of the pad); position j sits at (j + pre_lo) * up in the dilated array.
constexpr int kCtMaxPre = 8;

__device__ __forceinline__ int ct_preimages(int n, int s, int lo, int hi, int mode, int* out) {
  int c = 0;
  out[c++] = n;
  if (mode == PDEQX_PAD_ZEROS) return c;
  for (int q = -lo; q < 0; ++q)
    if (ct_boundary_map(q, s, mode) == n && c < kCtMaxPre) out[c++] = q;
  for (int q = s; q < s + hi; ++q)
    if (ct_boundary_map(q, s, mode) == n && c < kCtMaxPre) out[c++] = q;
  return c;
}

__global__ void __launch_bounds__(256)
convT_bwd_data_kernel(ConvTGeom g, const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ gx) {
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  const int64_t total = (int64_t)g.batch * g.c_in * g.n_in;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx % g.n_in;
    int64_t r = idx / g.n_in;
    const int ci = (int)(r % g.c_in);
    const int b = (int)(r / g.c_in);
    const int grp = ci / cig, il = ci % cig;
    int pre[PDEQX_MAX_DIMS][kCtMaxPre];
    int npre[PDEQX_MAX_DIMS] = {1, 1, 1};
    {
      int64_t nn = n;
      for (int a = g.ndim - 1; a >= 0; --a) {
        const int na = (int)(nn % g.spatial[a]);
        nn /= g.spatial[a];
        npre[a] = ct_preimages(na, g.spatial[a], g.pre_lo[a], g.pre_hi[a], g.boundary, pre[a]);
      }
    }
    const float* gyb = gy + ((int64_t)b * g.c_out + (int64_t)grp * cog) * g.n_out;
    const float* wg = w + ((int64_t)grp * cog * cig + il) * g.taps;
    float acc = 0.f;
    for (int c0 = 0; c0 < npre[0]; ++c0)
      for (int c1 = 0; c1 < npre[1]; ++c1)
        for (int c2 = 0; c2 < npre[2]; ++c2) {
          const int cc[PDEQX_MAX_DIMS] = {c0, c1, c2};
          for (int t = 0; t < g.taps; ++t) {
            int tt = t;
            int64_t dst = 0;
            bool valid = true;
            for (int a = g.ndim - 1; a >= 0; --a) {
              const int ta = tt % g.kernel[a];
              tt /= g.kernel[a];
              const int q = (pre[a][cc[a]] + g.pre_lo[a]) * g.up[a] - ta * g.dilation[a] + g.post_lo[a];
              if (q < 0 || q >= g.out[a]) { valid = false; break; }
              dst += (int64_t)q * g.out_stride[a];
            }
            if (!valid) continue;
#pragma unroll 8
            for (int o = 0; o < cog; ++o)
              acc = fmaf(__ldg(wg + (int64_t)o * cig * g.taps + t), __ldg(gyb + (int64_t)o * g.n_out + dst), acc);
          }
        }
    gx[idx] = acc;
  }
}

// filter gradient: CTA = 16 out-channels x 16 in-channels x one tap, reduction over (batch, output points)
constexpr int CT_WF_P = 64;

__global__ void __launch_bounds__(256)
convT_bwd_filter_kernel(ConvTGeom g, const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw,
                        float* __restrict__ gb, int o_tiles, int i_tiles, int64_t chunks) {
  __shared__ float Gs[CT_WF_P][16 + 1];
  __shared__ float Xs[CT_WF_P][16 + 1];
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  int bid = blockIdx.x;
  const int t = bid % g.taps;
  bid /= g.taps;
  const int it = bid % i_tiles;
  bid /= i_tiles;
  const int ot = bid % o_tiles;
  const int grp = bid / o_tiles;
  const int tid = threadIdx.x;
  const int ol = tid / 16, il = tid % 16;
  int tap[PDEQX_MAX_DIMS];
  {
    int tt = t;
    for (int a = g.ndim - 1; a >= 0; --a) {
      tap[a] = tt % g.kernel[a];
      tt /= g.kernel[a];
    }
  }
  float acc = 0.f, bsum = 0.f;
  // the reduction runs over the flattened (sample, output point) index (see conv_bwd_filter_kernel)
  const int64_t red = (int64_t)g.batch * g.n_out;
  (void)chunks;
  for (int64_t c = blockIdx.y; c * CT_WF_P < red; c += gridDim.y) {
    for (int idx = tid; idx < 16 * CT_WF_P; idx += 256) {
      const int ch = idx / CT_WF_P, pl = idx % CT_WF_P;
      const int64_t rp = c * CT_WF_P + pl;
      const int b = (int)(rp / g.n_out);
      const int64_t p = rp % g.n_out;
      float gv = 0.f, xv = 0.f;
      if (rp < red) {
        const int o = ot * 16 + ch, i = it * 16 + ch;
        if (o < cog) gv = __ldg(gy + ((int64_t)b * g.c_out + grp * cog + o) * g.n_out + p);
        if (i < cig) {
          int64_t pp = p, src = 0;
          bool valid = true;
          for (int a = g.ndim - 1; a >= 0; --a) {
            const int pa = (int)(pp % g.out[a]);
            pp /= g.out[a];
            const int m = ct_source(g, a, pa, tap[a]);
            valid = valid && (m >= 0);
            src += (int64_t)m * g.in_stride[a];
          }
          if (valid) xv = __ldg(x + ((int64_t)b * g.c_in + grp * cig + i) * g.n_in + src);
        }
      }
      Gs[pl][ch] = gv;
      Xs[pl][ch] = xv;
    }
    __syncthreads();
#pragma unroll 16
    for (int k = 0; k < CT_WF_P; ++k) {
      acc = fmaf(Gs[k][ol], Xs[k][il], acc);
      if (il == 0) bsum += Gs[k][ol];
    }
    __syncthreads();
  }
  const int o = ot * 16 + ol, i = it * 16 + il;
  if (o < cog && i < cig) atomicAdd(gw + ((int64_t)(grp * cog + o) * cig + i) * g.taps + t, acc);
  if (gb && il == 0 && it == 0 && t == 0 && o < cog) atomicAdd(gb + grp * cog + o, bsum);
}

static int convT_geometry(const pdeqx_convT_desc* d, ConvTGeom* out) {
  PDEQX_CHECK_ARG(d && out, "null argument");
  PDEQX_CHECK_ARG(d->ndim >= 1 && d->ndim <= PDEQX_MAX_DIMS, "ndim must be 1..3, got %d", d->ndim);
  PDEQX_CHECK_ARG(d->batch >= 0 && d->c_in >= 1 && d->c_out >= 1 && d->groups >= 1, "bad batch/channels/groups");
  PDEQX_CHECK_ARG(d->c_in % d->groups == 0 && d->c_out % d->groups == 0, "channels must be divisible by groups (=%d)", d->groups);
  PDEQX_CHECK_ARG(d->boundary >= PDEQX_PAD_ZEROS && d->boundary <= PDEQX_PAD_EDGE, "unknown boundary %d", d->boundary);
  ConvTGeom g{};
  g.ndim = d->ndim;
  g.batch = d->batch;
  g.c_in = d->c_in;
  g.c_out = d->c_out;
  g.groups = d->groups;
  g.boundary = d->boundary;
  g.taps = 1;
  g.n_in = 1;
  g.n_out = 1;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    const bool on = a < d->ndim;
    g.spatial[a] = on ? d->spatial[a] : 1;
    g.kernel[a] = on ? d->kernel[a] : 1;
    g.up[a] = on ? d->stride[a] : 1;
    g.dilation[a] = on ? d->dilation[a] : 1;
    PDEQX_CHECK_ARG(g.spatial[a] >= 1 && g.kernel[a] >= 1 && g.up[a] >= 1 && g.dilation[a] >= 1,
                    "spatial/kernel/stride/dilation must be >= 1 (axis %d)", a);
    const int p0 = on ? d->pad_lo[a] : 0, p1 = on ? d->pad_hi[a] : 0, op = on ? d->output_padding[a] : 0;
    PDEQX_CHECK_ARG(p0 >= 0 && p1 >= 0 && op >= 0, "negative padding (axis %d)", a);
    PDEQX_CHECK_ARG(op < g.up[a] || op == 0, "Must have `output_padding < stride` (elementwise).");
    int post_hi;
    if (d->boundary == PDEQX_PAD_ZEROS) {  // _conv.py:448-455
      g.pre_lo[a] = g.pre_hi[a] = 0;
      g.post_lo[a] = p0;
      post_hi = p1 + op;
    } else {  // _conv.py:411-436
      const int tp0 = g.dilation[a] * (g.kernel[a] - 1) - p0, tp1 = g.dilation[a] * (g.kernel[a] - 1) - p1 + op;
      PDEQX_CHECK_ARG(tp0 >= 0 && tp1 >= 0, "padding larger than the dilated kernel extent (axis %d)", a);
      g.pre_lo[a] = (tp0 + g.up[a] - 1) / g.up[a];
      g.pre_hi[a] = (tp1 + g.up[a] - 1) / g.up[a];
      g.post_lo[a] = p0 - g.pre_lo[a] * g.up[a];
      post_hi = p1 - g.pre_hi[a] * g.up[a] + op;
      if (d->boundary == PDEQX_PAD_REFLECT)
        PDEQX_CHECK_ARG(g.pre_lo[a] < g.spatial[a] && g.pre_hi[a] < g.spatial[a], "reflect padding needs more points on axis %d", a);
    }
    g.vlen[a] = (g.spatial[a] + g.pre_lo[a] + g.pre_hi[a] - 1) * g.up[a] + 1;
    const int span = g.vlen[a] + g.post_lo[a] + post_hi - g.dilation[a] * (g.kernel[a] - 1);
    PDEQX_CHECK_ARG(span >= 1, "kernel does not fit the padded input on axis %d", a);
    g.out[a] = span;
    g.taps *= g.kernel[a];
    g.n_in *= g.spatial[a];
    g.n_out *= g.out[a];
  }
  for (int a = PDEQX_MAX_DIMS - 1, si = 1, so = 1; a >= 0; --a) {
    g.in_stride[a] = si;
    g.out_stride[a] = so;
    si *= g.spatial[a];
    so *= g.out[a];
  }
  *out = g;
  return PDEQX_OK;
}

static inline int ct_grid_for(int64_t total) {
  int64_t b = ceil_div(total, 256);
  const int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace pdeqx

using namespace pdeqx;

extern "C" int pdeqx_convT_out_spatial(const pdeqx_convT_desc* d, int32_t out_spatial[PDEQX_MAX_DIMS]) {
  ConvTGeom g;
  PDEQX_TRY(convT_geometry(d, &g));
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) out_spatial[a] = a < d->ndim ? g.out[a] : 1;
  return PDEQX_OK;
}

extern "C" int pdeqx_convT_fwd(const pdeqx_convT_desc* d, const float* x, const float* w, const float* b, float* y,
                               void* stream) {
  PDEQX_CHECK_ARG(x && w && y, "null argument");
  ConvTGeom g;
  PDEQX_TRY(convT_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  const int cog = g.c_out / g.groups;
  const int64_t total = (int64_t)g.batch * g.groups * ceil_div(cog, kCtOPT) * g.n_out;
  if (g.c_in / g.groups >= 32 && g.n_in <= 1024 && g.n_out <= 2048)
    convT_fwd_kernel<true><<<ct_grid_for(total * 32), 256, 0, (cudaStream_t)stream>>>(g, x, w, b, y);
  else
    convT_fwd_kernel<false><<<ct_grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, x, w, b, y);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_convT_bwd_data(const pdeqx_convT_desc* d, const float* gy, const float* w, float* gx, void* stream) {
  PDEQX_CHECK_ARG(gy && w && gx, "null argument");
  ConvTGeom g;
  PDEQX_TRY(convT_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  for (int a = 0; a < g.ndim; ++a) {
    int worst = 1;
    const int s = g.spatial[a], lo = g.pre_lo[a], hi = g.pre_hi[a];
    if (g.boundary == PDEQX_PAD_WRAP) worst = 1 + (lo + s - 1) / s + (hi + s - 1) / s;
    else if (g.boundary == PDEQX_PAD_REFLECT) worst = 3;
    else if (g.boundary == PDEQX_PAD_EDGE) worst = 1 + (lo > hi ? lo : hi);
    if (worst > kCtMaxPre) {
      set_error("padding maps onto one point %d times (> %d): grid too small for this kernel", worst, kCtMaxPre);
      return PDEQX_ERR_UNSUPPORTED;
    }
  }
  const int64_t total = (int64_t)g.batch * g.c_in * g.n_in;
  convT_bwd_data_kernel<<<ct_grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, gy, w, gx);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_convT_bwd_filter(const pdeqx_convT_desc* d, const float* x, const float* gy, float* gw, float* gb,
                                      void* stream) {
  PDEQX_CHECK_ARG(x && gy && gw, "null argument");
  ConvTGeom g;
  PDEQX_TRY(convT_geometry(d, &g));
  cudaStream_t s = (cudaStream_t)stream;
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  // (a kernel node, not two memset nodes: inside a replayed graph a memset node costs ~3 us of idle time, see core.cu)
  PDEQX_TRY(zero_small(gw, (int64_t)g.c_out * cig * g.taps, gb, g.c_out, s));
  if (g.batch == 0) return PDEQX_OK;
  const int o_tiles = (int)ceil_div(cog, 16), i_tiles = (int)ceil_div(cig, 16);
  const int64_t chunks = ceil_div(g.n_out, CT_WF_P);
  const int64_t gx_blocks = (int64_t)g.groups * o_tiles * i_tiles * g.taps;
  PDEQX_CHECK_ARG(gx_blocks < (1ll << 31), "convT_bwd_filter: grid too large");
  int64_t split = ceil_div(148 * 8, gx_blocks);
  if (split > ceil_div((int64_t)g.batch * g.n_out, CT_WF_P)) split = ceil_div((int64_t)g.batch * g.n_out, CT_WF_P);
  if (split > 65535) split = 65535;
  if (split < 1) split = 1;
  convT_bwd_filter_kernel<<<dim3((unsigned)gx_blocks, (unsigned)split), 256, 0, s>>>(g, x, gy, gw, gb, o_tiles, i_tiles, chunks);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

// channel concatenation of the UNet skip connection (jnp.concatenate((skip, x), axis=0), _hierarchical.py:275):
// out[B, ca + cb, n] = [a[B, ca, n] ; b[B, cb, n]] -- two strided device copies on `stream`
extern "C" int pdeqx_concat_channels(int32_t batch, int32_t c_a, int32_t c_b, int64_t n, const float* a, const float* b,
                                     float* out, void* stream) {
  PDEQX_CHECK_ARG(a && b && out && batch >= 0 && c_a >= 1 && c_b >= 1 && n >= 0, "bad arguments");
  if (batch == 0 || n == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t row_a = (size_t)c_a * n * 4, row_b = (size_t)c_b * n * 4, row_o = row_a + row_b;
  PDEQX_CUDA(cudaMemcpy2DAsync(out, row_o, a, row_a, row_a, batch, cudaMemcpyDeviceToDevice, s));
  PDEQX_CUDA(cudaMemcpy2DAsync((char*)out + row_a, row_o, b, row_b, row_b, batch, cudaMemcpyDeviceToDevice, s));
  return PDEQX_OK;
}

// adjoint: ga = g[:, :c_a] (+ add_a), gb = g[:, c_a:]; add_a (the other cotangent reaching the skip tensor) may be NULL
extern "C" int pdeqx_split_channels(int32_t batch, int32_t c_a, int32_t c_b, int64_t n, const float* g, float* ga, float* gb,
                                    void* stream) {
  PDEQX_CHECK_ARG(g && ga && gb && batch >= 0 && c_a >= 1 && c_b >= 1 && n >= 0, "bad arguments");
  if (batch == 0 || n == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t row_a = (size_t)c_a * n * 4, row_b = (size_t)c_b * n * 4, row_o = row_a + row_b;
  PDEQX_CUDA(cudaMemcpy2DAsync(ga, row_a, g, row_o, row_a, batch, cudaMemcpyDeviceToDevice, s));
  PDEQX_CUDA(cudaMemcpy2DAsync(gb, row_b, (const char*)g + row_a, row_o, row_b, batch, cudaMemcpyDeviceToDevice, s));
  return PDEQX_OK;
}
