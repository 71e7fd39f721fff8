// PhysicsConv / MorePaddingConv: boundary-aware N-D cross-correlation, generic fp32 path.
//
//   replaces pdequinox/conv/_conv.py:169-219 (jnp.pad + lax.conv_general_dilated + bias) with the
//   SAME padding of pdequinox/conv/_physics_conv.py:11-26.
//
// The halo is never materialised: every tap computes its source index through the boundary map
//   ZEROS   (dirichlet) : out-of-range taps contribute 0          (_conv.py:192-194)
//   WRAP    (periodic)  : q mod s                                 (_conv.py:190, jnp.pad "wrap")
//   REFLECT (neumann)   : mirror without repeating the edge       (_conv.py:196, jnp.pad "reflect")
//   EDGE    (replicate) : clamp                                   (_conv.py:199)
// This file is the any-shape path (1-3 D, stride, dilation, groups, even kernels); the C >= 16
// 2-D stride-1 case runs on tcgen05 as an implicit GEMM (tc_conv.cu).
#include "common.cuh"
#include "conv_plan.h"

namespace pdeqx {

__host__ __device__ __forceinline__ int boundary_map(int q, int s, int mode) {
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
      return -1;  // zeros
  }
}

// ---------------------------------------------------------------------------------------------
// forward: one thread per output element, 4 output channels per thread
// ---------------------------------------------------------------------------------------------
constexpr int kOPT = 4;

__global__ void __launch_bounds__(256)
conv_fwd_kernel(ConvGeom g, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ residual, float* __restrict__ y, int act) {
  const int64_t n_out = g.n_out;
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  const int ochunks = (cog + kOPT - 1) / kOPT;  // per group
  const int64_t total = (int64_t)g.batch * g.groups * ochunks * n_out;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = idx % n_out;
    int64_t r = idx / n_out;
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
    const int o0 = grp * cog + oc * kOPT;
    const int no = min(kOPT, cog - oc * kOPT);
    float acc[kOPT];
#pragma unroll
    for (int j = 0; j < kOPT; ++j) acc[j] = 0.f;
    const float* xb = x + ((int64_t)b * g.c_in + (int64_t)grp * cig) * g.n_in;
    for (int t = 0; t < g.taps; ++t) {
      int tt = t;
      int64_t src = 0;
      bool valid = true;
      for (int a = g.ndim - 1; a >= 0; --a) {
        int ta = tt % g.kernel[a];
        tt /= g.kernel[a];
        int q = po[a] * g.stride[a] + ta * g.dilation[a] - g.pad_lo[a];
        int m = boundary_map(q, g.spatial[a], g.boundary);
        valid = valid && (m >= 0);
        src += (int64_t)m * g.in_stride[a];
      }
      if (!valid) continue;
#pragma unroll 8
      for (int i = 0; i < cig; ++i) {  // independent loads in flight: the small-tensor regime is latency bound
        float xv = __ldg(xb + (int64_t)i * g.n_in + src);
#pragma unroll
        for (int j = 0; j < kOPT; ++j)
          if (j < no) acc[j] = fmaf(__ldg(w + ((int64_t)(o0 + j) * cig + i) * g.taps + t), xv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kOPT; ++j) {
      if (j >= no) break;
      int64_t off = ((int64_t)b * g.c_out + o0 + j) * n_out + p;
      float v = acc[j] + (bias ? __ldg(bias + o0 + j) : 0.f);
      if (residual) v += __ldg(residual + off);
      y[off] = apply_act_rt(v, act);
    }
  }
}

// Deep-channel / few-points regime (the inner levels of ClassicUNet 1D: 256 channels on 4 points): one thread per output
// would loop serially over c_in/groups x taps = 768 dependent loads (~50 us per launch on KB-sized tensors). Here ONE WARP
// owns an output quad and its lanes split the input channels; a shuffle reduction finishes. Everything is cache resident, so
// the lane-strided x reads do not matter.
__global__ void __launch_bounds__(256)
conv_fwd_warp_kernel(ConvGeom g, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ residual, float* __restrict__ y, int act) {
  const int64_t n_out = g.n_out;
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  const int ochunks = (cog + kOPT - 1) / kOPT;
  const int64_t total = (int64_t)g.batch * g.groups * ochunks * n_out;
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < total; idx += warps) {
    int64_t p = idx % n_out;
    int64_t r = idx / n_out;
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
    const int o0 = grp * cog + oc * kOPT;
    const int no = min(kOPT, cog - oc * kOPT);
    float acc[kOPT];
#pragma unroll
    for (int j = 0; j < kOPT; ++j) acc[j] = 0.f;
    const float* xb = x + ((int64_t)b * g.c_in + (int64_t)grp * cig) * g.n_in;
    for (int t = 0; t < g.taps; ++t) {
      int tt = t;
      int64_t src = 0;
      bool valid = true;
      for (int a = g.ndim - 1; a >= 0; --a) {
        int ta = tt % g.kernel[a];
        tt /= g.kernel[a];
        int q = po[a] * g.stride[a] + ta * g.dilation[a] - g.pad_lo[a];
        int m = boundary_map(q, g.spatial[a], g.boundary);
        valid = valid && (m >= 0);
        src += (int64_t)m * g.in_stride[a];
      }
      if (!valid) continue;
      for (int i = lane; i < cig; i += 32) {
        const float xv = __ldg(xb + (int64_t)i * g.n_in + src);
#pragma unroll
        for (int j = 0; j < kOPT; ++j)
          if (j < no) acc[j] = fmaf(__ldg(w + ((int64_t)(o0 + j) * cig + i) * g.taps + t), xv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kOPT; ++j)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    if (lane < no) {
      float v = acc[0];
#pragma unroll
      for (int j = 1; j < kOPT; ++j)
        if (lane == j) v = acc[j];
      const int64_t off = ((int64_t)b * g.c_out + o0 + lane) * n_out + p;
      v += bias ? __ldg(bias + o0 + lane) : 0.f;
      if (residual) v += __ldg(residual + off);
      y[off] = apply_act_rt(v, act);
    }
  }
}

// true when the warp-cooperative kernels pay: enough reduction work per output to fill a warp, tensors small enough to be
// cache resident (lanes read channel-strided)
static inline bool conv_small_deep(const ConvGeom& g, int red_channels) {
  return red_channels >= 32 && g.n_in <= 1024 && g.n_out <= 1024;
}

// ---------------------------------------------------------------------------------------------
// backward data: gather over the pre-images of each input position under the boundary map
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPre = 8;

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ int preimages(int n, int s, int lo, int hi, int mode, int* out) {
  int c = 0;
  out[c++] = n;
  if (mode == PDEQX_PAD_ZEROS) return c;
  for (int q = -lo; q < 0; ++q)
    if (boundary_map(q, s, mode) == n && c < kMaxPre) out[c++] = q;
  for (int q = s; q < s + hi; ++q)
    if (boundary_map(q, s, mode) == n && c < kMaxPre) out[c++] = q;
  return c;
}

// MIRROR_ONLY: only the contributions that reach an input point through a wrapped / reflected halo position, ADDED to gx
// (the tcgen05 path computes the zero-extended part; see tc_conv_bwd_data for neumann)
// (2-D only) the points of one channel plane within `band` of an edge, enumerated as top rows, bottom rows, then the left and
// right columns of the rows in between: only these can be reached through a mirrored / wrapped halo position
struct Frame2d {
  int band_h, band_w;  // band widths along axis 0 / axis 1 (clipped so that the bands do not overlap)
  int64_t count;       // points per channel plane
};
__host__ __device__ __forceinline__ Frame2d frame2d(const ConvGeom& g) {
  Frame2d f;
  const int H = g.spatial[0], W = g.spatial[1];
  const int bh = g.pad_lo[0] > g.pad_hi[0] ? g.pad_lo[0] : g.pad_hi[0], bw = g.pad_lo[1] > g.pad_hi[1] ? g.pad_lo[1] : g.pad_hi[1];
  f.band_h = 2 * (bh + 1) >= H ? (H + 1) / 2 : bh + 1;  // reflect reaches index `pad`, wrap index pad - 1
  f.band_w = 2 * (bw + 1) >= W ? (W + 1) / 2 : bw + 1;
  const int mid_h = H - 2 * f.band_h > 0 ? H - 2 * f.band_h : 0;
  const int top = H - mid_h;  // rows covered entirely
  const int side = W - 2 * f.band_w > 0 ? 2 * f.band_w : W;
  f.count = (int64_t)top * W + (int64_t)mid_h * side;
  return f;
}
__device__ __forceinline__ int64_t frame2d_point(const ConvGeom& g, const Frame2d& f, int64_t k) {
  const int H = g.spatial[0], W = g.spatial[1];
  const int mid_h = H - 2 * f.band_h > 0 ? H - 2 * f.band_h : 0;
  const int top = H - mid_h;
  if (k < (int64_t)top * W) {  // full rows: the first band_h rows, then the last ones
    const int r = (int)(k / W), c = (int)(k % W);
    const int h = r < f.band_h ? r : r + mid_h;
    return (int64_t)h * W + c;
  }
  k -= (int64_t)top * W;
  const int side = W - 2 * f.band_w > 0 ? 2 * f.band_w : W;
  const int r = (int)(k / side), c = (int)(k % side);
  const int wcol = (side == W || c < f.band_w) ? c : W - side + c;
  return (int64_t)(f.band_h + r) * W + wcol;
}

template <bool MIRROR_ONLY>
__global__ void __launch_bounds__(256)
conv_bwd_data_kernel(ConvGeom g, const float* __restrict__ gy, const float* __restrict__ w, const float* __restrict__ add,
                     float* __restrict__ gx) {
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  const bool framed = MIRROR_ONLY && g.ndim == 2;  // walk only the frame of every plane
  Frame2d fr{};
  if (framed) fr = frame2d(g);
  const int64_t per_plane = framed ? fr.count : g.n_in;
  const int64_t total = (int64_t)g.batch * g.c_in * per_plane;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = it % per_plane;
    int64_t r = it / per_plane;
    if (framed) n = frame2d_point(g, fr, n);
    const int64_t idx = r * g.n_in + n;
    const int ci = (int)(r % g.c_in);
    const int b = (int)(r / g.c_in);
    const int grp = ci / cig, il = ci % cig;
    int pre[PDEQX_MAX_DIMS][kMaxPre];
    int npre[PDEQX_MAX_DIMS] = {1, 1, 1};
    {
      int64_t nn = n;
      for (int a = g.ndim - 1; a >= 0; --a) {
        int na = (int)(nn % g.spatial[a]);
        nn /= g.spatial[a];
        npre[a] = preimages(na, g.spatial[a], g.pad_lo[a], g.pad_hi[a], g.boundary, pre[a]);
      }
    }
    if (MIRROR_ONLY && npre[0] * npre[1] * npre[2] == 1) continue;
    const float* gyb = gy + ((int64_t)b * g.c_out + (int64_t)grp * cog) * g.n_out;
    const float* wg = w + ((int64_t)grp * cog * cig + il) * g.taps;
    float acc = 0.f;
    for (int c0 = 0; c0 < npre[0]; ++c0)
      for (int c1 = 0; c1 < npre[1]; ++c1)
        for (int c2 = 0; c2 < npre[2]; ++c2) {
          if (MIRROR_ONLY && c0 + c1 + c2 == 0) continue;  // the direct position is the zero-extended part
          const int cc[PDEQX_MAX_DIMS] = {c0, c1, c2};
          for (int t = 0; t < g.taps; ++t) {
            int tt = t;
            int64_t dst = 0;
            bool valid = true;
            for (int a = g.ndim - 1; a >= 0; --a) {
              int ta = tt % g.kernel[a];
              tt /= g.kernel[a];
              int num = pre[a][cc[a]] + g.pad_lo[a] - ta * g.dilation[a];
              if (num < 0 || num % g.stride[a] != 0) { valid = false; break; }
              int pa = num / g.stride[a];
              if (pa >= g.out[a]) { valid = false; break; }
              dst += (int64_t)pa * g.out_stride[a];
            }
            if (!valid) continue;
#pragma unroll 8
            for (int o = 0; o < cog; ++o)
              acc = fmaf(__ldg(wg + (int64_t)o * cig * g.taps + t), __ldg(gyb + (int64_t)o * g.n_out + dst), acc);
          }
        }
    if (MIRROR_ONLY) gx[idx] += acc;
    else gx[idx] = acc + (add ? __ldg(add + idx) : 0.f);
  }
}


// Mirror / wrap part of the data gradient for the 2-D, groups == 1, c_in == c_out == 64 case with the whole filter
// resident in shared memory (the tcgen05 path: 64 x 64 x 9 fp32 = 147 KB). Persistent blocks of 32 warps; ONE WARP per
// frame point (32 points in flight per SM hide the load latency; lanes own channels c and c + 32). For its point a warp
// enumerates the (pre-image, tap) pairs that hit an output position, first fetches gy[b, o, position] of every pair into
// its own shared-memory slab -- all of the point's global loads in flight together -- then contracts them with the filter
// (conflict-free filter reads, broadcast gy reads). Only __syncwarp between the phases.
//   gx[b,ci,n] += sum over the non-direct pre-images q of n under the boundary map, taps t, out channels o:
//                 w[o,ci,t] * gy[b,o,(q + pad - t d) / stride]
constexpr int kMirPairs = 8;    // pairs staged per batch
constexpr int kMirWarps = 32;
__global__ void __launch_bounds__(kMirWarps * 32)
conv_bwd_data_mirror2d_kernel(ConvGeom g, const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ gx) {
  extern __shared__ float ws[];  // [taps][64 o][64 ci], then [warps][kMirPairs][64 o]
  constexpr int CH = 64;
  const int taps = g.taps;
  float* gys = ws + (size_t)taps * CH * CH;
  for (int idx = threadIdx.x; idx < taps * CH * CH; idx += blockDim.x) {
    const int i = idx % CH, o = (idx / CH) % CH, t = idx / (CH * CH);
    ws[idx] = __ldg(w + ((int64_t)o * CH + i) * taps + t);
  }
  __syncthreads();
  const Frame2d fr = frame2d(g);
  const int warp = threadIdx.x >> 5, c = threadIdx.x & 31;
  float* mine = gys + (size_t)warp * kMirPairs * CH;
  const int64_t total = (int64_t)g.batch * fr.count;
  for (int64_t pt = (int64_t)blockIdx.x * kMirWarps + warp; pt < total; pt += (int64_t)gridDim.x * kMirWarps) {
    const int b = (int)(pt / fr.count);
    const int64_t n = frame2d_point(g, fr, pt % fr.count);
    int pre[2][kMaxPre];
    int npre[2];
    npre[1] = preimages((int)(n % g.spatial[1]), g.spatial[1], g.pad_lo[1], g.pad_hi[1], g.boundary, pre[1]);
    npre[0] = preimages((int)(n / g.spatial[1]), g.spatial[0], g.pad_lo[0], g.pad_hi[0], g.boundary, pre[0]);
    if (npre[0] * npre[1] == 1) continue;
    const float* gyc = gy + ((int64_t)b * CH + c) * g.n_out;  // this lane's output channels c, c + 32 in the fetch phase
    float* gxp = gx + ((int64_t)b * CH + c) * g.n_in + n;
    const float old0 = gxp[0], old1 = gxp[(int64_t)32 * g.n_in];  // (in flight while the pairs are processed)
    float acc0 = 0.f, acc1 = 0.f;
    const int npairs_all = npre[0] * npre[1] * taps;  // pair index k = (c0 * npre[1] + c1) * taps + t
    int k = taps;                                       // skip (c0, c1) = (0, 0): the direct position
    while (k < npairs_all) {
      // ---- fetch phase: up to kMirPairs valid pairs ----
      int tl[kMirPairs];
      int cnt = 0;
      for (; k < npairs_all && cnt < kMirPairs; ++k) {
        const int t = k % taps, cc = k / taps, c1 = cc % npre[1], c0 = cc / npre[1];
        const int t0 = t / g.kernel[1], t1 = t % g.kernel[1];
        const int num0 = pre[0][c0] + g.pad_lo[0] - t0 * g.dilation[0], num1 = pre[1][c1] + g.pad_lo[1] - t1 * g.dilation[1];
        if (num0 < 0 || num1 < 0 || num0 % g.stride[0] != 0 || num1 % g.stride[1] != 0) continue;
        const int p0 = num0 / g.stride[0], p1 = num1 / g.stride[1];
        if (p0 >= g.out[0] || p1 >= g.out[1]) continue;
        const float* src = gyc + (int64_t)p0 * g.out_stride[0] + p1;
        mine[cnt * CH + c] = __ldg(src);
        mine[cnt * CH + c + 32] = __ldg(src + (int64_t)32 * g.n_out);
        tl[cnt++] = t;
      }
      __syncwarp();
      // ---- contraction phase ----
      for (int j = 0; j < cnt; ++j) {
        const float* wt = ws + (size_t)tl[j] * CH * CH + c;
        const float* gv = mine + j * CH;
#pragma unroll 16
        for (int o = 0; o < CH; ++o) {
          const float gvo = gv[o];
          acc0 = fmaf(wt[o * CH], gvo, acc0);
          acc1 = fmaf(wt[o * CH + 32], gvo, acc1);
        }
      }
      __syncwarp();
    }
    gxp[0] = old0 + acc0;
    gxp[(int64_t)32 * g.n_in] = old1 + acc1;
  }
}

// ---------------------------------------------------------------------------------------------
// backward filter: CTA = 16 out-channels x 16 in-channels x one tap, reduction over (batch, pixels)
// ---------------------------------------------------------------------------------------------
constexpr int WF_P = 64;

__global__ void __launch_bounds__(256)
conv_bwd_filter_kernel(ConvGeom g, const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw,
                       float* __restrict__ gb, int o_tiles, int i_tiles, int64_t chunks) {
  __shared__ float Gs[WF_P][16 + 1];
  __shared__ float Xs[WF_P][16 + 1];
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
  // the reduction runs over the flattened (sample, pixel) index in chunks of WF_P: with a handful of points per sample (the
  // inner UNet levels) a per-sample chunk would be almost all padding
  const int64_t red = (int64_t)g.batch * g.n_out;
  (void)chunks;
  for (int64_t c = blockIdx.y; c * WF_P < red; c += gridDim.y) {
    for (int idx = tid; idx < 16 * WF_P; idx += 256) {
      const int ch = idx / WF_P, pl = idx % WF_P;
      const int64_t rp = c * WF_P + pl;
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
            int pa = (int)(pp % g.out[a]);
            pp /= g.out[a];
            int m = boundary_map(pa * g.stride[a] + tap[a] * g.dilation[a] - g.pad_lo[a], g.spatial[a], g.boundary);
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
    for (int k = 0; k < WF_P; ++k) {
      acc = fmaf(Gs[k][ol], Xs[k][il], acc);
      if (il == 0) bsum += Gs[k][ol];
    }
    __syncthreads();
  }
  const int o = ot * 16 + ol, i = it * 16 + il;
  if (o < cog && i < cig) atomicAdd(gw + ((int64_t)(grp * cog + o) * cig + i) * g.taps + t, acc);
  if (gb && il == 0 && it == 0 && t == 0 && o < cog) atomicAdd(gb + grp * cog + o, bsum);
}

int conv_geometry(const pdeqx_conv_desc* d, ConvGeom* out) {
  PDEQX_CHECK_ARG(d && out, "null argument");
  PDEQX_CHECK_ARG(d->ndim >= 1 && d->ndim <= PDEQX_MAX_DIMS, "ndim must be 1..3, got %d", d->ndim);
  PDEQX_CHECK_ARG(d->batch >= 0 && d->c_in >= 1 && d->c_out >= 1 && d->groups >= 1, "bad batch/channels/groups");
  PDEQX_CHECK_ARG(d->c_in % d->groups == 0, "`in_channels` (=%d) must be divisible by `groups` (=%d).", d->c_in, d->groups);
  PDEQX_CHECK_ARG(d->c_out % d->groups == 0, "out_channels (=%d) must be divisible by groups (=%d)", d->c_out, d->groups);
  PDEQX_CHECK_ARG(d->boundary >= PDEQX_PAD_ZEROS && d->boundary <= PDEQX_PAD_EDGE, "unknown boundary %d", d->boundary);
  PDEQX_CHECK_ARG(d->precision == PDEQX_PREC_FP32 || d->precision == PDEQX_PREC_TF32, "unknown precision %d", d->precision);
  ConvGeom g{};
  g.ndim = d->ndim;
  g.batch = d->batch;
  g.c_in = d->c_in;
  g.c_out = d->c_out;
  g.groups = d->groups;
  g.boundary = d->boundary;
  g.precision = d->precision;
  g.taps = 1;
  g.n_in = 1;
  g.n_out = 1;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    const bool on = a < d->ndim;
    g.spatial[a] = on ? d->spatial[a] : 1;
    g.kernel[a] = on ? d->kernel[a] : 1;
    g.stride[a] = on ? d->stride[a] : 1;
    g.dilation[a] = on ? d->dilation[a] : 1;
    g.pad_lo[a] = on ? d->pad_lo[a] : 0;
    g.pad_hi[a] = on ? d->pad_hi[a] : 0;
    PDEQX_CHECK_ARG(g.spatial[a] >= 1 && g.kernel[a] >= 1 && g.stride[a] >= 1 && g.dilation[a] >= 1,
                    "spatial/kernel/stride/dilation must be >= 1 (axis %d)", a);
    PDEQX_CHECK_ARG(g.pad_lo[a] >= 0 && g.pad_hi[a] >= 0, "negative padding (axis %d)", a);
    if (d->boundary == PDEQX_PAD_REFLECT && on)
      PDEQX_CHECK_ARG(g.spatial[a] > 1 || (g.pad_lo[a] == 0 && g.pad_hi[a] == 0),
                      "reflect padding needs at least 2 points on axis %d", a);
    const int span = g.spatial[a] + g.pad_lo[a] + g.pad_hi[a] - g.dilation[a] * (g.kernel[a] - 1) - 1;
    PDEQX_CHECK_ARG(span >= 0, "kernel does not fit the padded input on axis %d", a);
    g.out[a] = span / g.stride[a] + 1;
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
  // pre-image multiplicity of the boundary map must fit the gather in the data-gradient kernel
  // (upper bound per axis: one hit per period of the map inside each halo)
  g.max_pre = 1;
  if (d->boundary != PDEQX_PAD_ZEROS) {
    for (int a = 0; a < d->ndim; ++a) {
      const int s = g.spatial[a], lo = g.pad_lo[a], hi = g.pad_hi[a];
      int worst = 1;
      if (d->boundary == PDEQX_PAD_WRAP) worst = 1 + (lo + s - 1) / s + (hi + s - 1) / s;
      else if (d->boundary == PDEQX_PAD_REFLECT) {
        const int per = 2 * (s - 1);  // both +n and -n occur once per period
        worst = s > 1 ? 1 + 2 * ((lo + per - 1) / per) + 2 * ((hi + per - 1) / per) : 1;
      }
      else worst = 1 + (lo > hi ? lo : hi);
      g.max_pre = worst > g.max_pre ? worst : g.max_pre;
    }
  }
  *out = g;
  return PDEQX_OK;
}

static inline int grid_for(int64_t total) {
  int64_t b = ceil_div(total, 256);
  const int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

int conv_bwd_data_mirror_part(const ConvGeom& g, const float* gy, const float* w, float* gx, cudaStream_t s) {
  const size_t wbytes = sizeof(float) * ((size_t)g.taps * g.c_out * g.c_in + kMirWarps * kMirPairs * 64);
  if (g.ndim == 2 && g.groups == 1 && g.c_in == 64 && g.c_out == 64 && wbytes <= 220 * 1024) {
    static DeviceOnce attr;      
    if (attr.first_use()) {
      PDEQX_CUDA(cudaFuncSetAttribute(conv_bwd_data_mirror2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                  
    }
    conv_bwd_data_mirror2d_kernel<<<148, kMirWarps * 32, wbytes, s>>>(g, gy, w, gx);
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
  const int64_t total = (int64_t)g.batch * g.c_in * (g.ndim == 2 ? frame2d(g).count : g.n_in);
  conv_bwd_data_kernel<true><<<grid_for(total), 256, 0, s>>>(g, gy, w, nullptr, gx);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx

using namespace pdeqx;

namespace pdeqx { int64_t tc_conv_pair_launches(); }
extern "C" int64_t pdeqx_conv_pair_launches(void) { return pdeqx::tc_conv_pair_launches(); }

extern "C" int pdeqx_conv_tcgen05_passes(const pdeqx_conv_desc* d) {
  ConvGeom g;
  if (conv_geometry(d, &g) != PDEQX_OK) return 0;
  return (tc_conv_available(g) ? 1 : 0) | (tc_conv_bwd_data_available(g) ? 2 : 0) | (tc_conv_bwd_filter_available(g) ? 4 : 0);
}

extern "C" size_t pdeqx_conv_workspace_bytes(const pdeqx_conv_desc* d) {
  ConvGeom g;
  if (conv_geometry(d, &g) != PDEQX_OK) return 0;
  return tc_conv_workspace_bytes(g);
}

extern "C" int pdeqx_conv_out_spatial(const pdeqx_conv_desc* d, int32_t out_spatial[PDEQX_MAX_DIMS]) {
  ConvGeom g;
  PDEQX_TRY(conv_geometry(d, &g));
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) out_spatial[a] = a < d->ndim ? g.out[a] : 1;
  return PDEQX_OK;
}

extern "C" int pdeqx_conv_fwd(const pdeqx_conv_desc* d, const float* x, const float* w, const float* b,
                              const float* residual, int32_t activation, float* y, void* workspace,
                              size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(x && w && y, "null argument");
  PDEQX_CHECK_ARG(activation >= 0 && activation <= 2, "unknown activation %d", activation);
  ConvGeom g;
  PDEQX_TRY(conv_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (tc_conv_available(g)) return tc_conv_fwd(g, x, w, b, residual, activation, y, workspace, workspace_bytes, s);
  const int cog = g.c_out / g.groups;
  const int64_t total = (int64_t)g.batch * g.groups * ceil_div(cog, kOPT) * g.n_out;
  TimedScope timed(PDEQX_K_CONV_FWD, s);
  if (conv_small_deep(g, g.c_in / g.groups)) conv_fwd_warp_kernel<<<grid_for(total * 32), 256, 0, s>>>(g, x, w, b, residual, y, activation);
  else conv_fwd_kernel<<<grid_for(total), 256, 0, s>>>(g, x, w, b, residual, y, activation);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

extern "C" int pdeqx_conv_bwd_data(const pdeqx_conv_desc* d, const float* gy, const float* w, const float* add,
                                   const float* relu_mask, float* gx, void* workspace, size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(gy && w && gx, "null argument");
  ConvGeom g;
  PDEQX_TRY(conv_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  if (g.max_pre > kMaxPre) {
    set_error("padding wraps/reflects onto one point %d times (> %d): grid too small for this dilation", g.max_pre, kMaxPre);
    return PDEQX_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t total = (int64_t)g.batch * g.c_in * g.n_in;
  bool masked = false;  // the tcgen05 kernels zero the masked elements in their epilogue; every other path in a pass behind
  if (tc_conv_bwd_data_available(g)) {
    PDEQX_TRY(tc_conv_bwd_data(g, gy, w, add, relu_mask, gx, workspace, workspace_bytes, s));
    masked = g.boundary != PDEQX_PAD_REFLECT;
  } else {
    // (a warp-per-point variant of this kernel, lanes over the output channels, was measured SLOWER -- 57-74 us vs 35 us per
    // launch on the UNet levels: the pre-image enumeration is replicated by every lane -- and removed)
    conv_bwd_data_kernel<false><<<grid_for(total), 256, 0, s>>>(g, gy, w, add, gx);
    PDEQX_LAUNCHED();
  }
  if (relu_mask && !masked) return pdeqx_activation_bwd(total, relu_mask, gx, PDEQX_ACT_RELU, gx, stream);
  return PDEQX_OK;
}

extern "C" int pdeqx_conv_bwd_filter(const pdeqx_conv_desc* d, const float* x, const float* gy, float* gw, float* gb,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(x && gy && gw, "null argument");
  ConvGeom g;
  PDEQX_TRY(conv_geometry(d, &g));
  cudaStream_t s = (cudaStream_t)stream;
  const int cog = g.c_out / g.groups, cig = g.c_in / g.groups;
  // (the tcgen05 pass overwrites every element of gw / gb: no memset nodes in front of it)
  if (g.batch > 0 && tc_conv_bwd_filter_available(g)) return tc_conv_bwd_filter(g, x, gy, gw, gb, workspace, workspace_bytes, s);
  // (a kernel node, not two memset nodes: inside a replayed graph a memset node costs ~3 us of idle time, see core.cu)
  PDEQX_TRY(zero_small(gw, (int64_t)g.c_out * cig * g.taps, gb, g.c_out, s));
  if (g.batch == 0) return PDEQX_OK;
  const int o_tiles = (int)ceil_div(cog, 16), i_tiles = (int)ceil_div(cig, 16);
  const int64_t chunks = ceil_div(g.n_out, WF_P);
  const int64_t gx_blocks = (int64_t)g.groups * o_tiles * i_tiles * g.taps;
  PDEQX_CHECK_ARG(gx_blocks < (1ll << 31), "conv_bwd_filter: grid too large");
  int64_t split = ceil_div(148 * 8, gx_blocks);
  const int64_t red_chunks = ceil_div((int64_t)g.batch * g.n_out, WF_P);  // chunks of the flattened (sample, pixel) reduction
  if (split > red_chunks) split = red_chunks;
  if (split > 65535) split = 65535;
  if (split < 1) split = 1;
  conv_bwd_filter_kernel<<<dim3((unsigned)gx_blocks, (unsigned)split), 256, 0, s>>>(g, x, gy, gw, gb, o_tiles, i_tiles, chunks);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

// ------------------------------------------------------------------------------------------
// GroupNorm (groups = 1) -> PhysicsConv -> activation, the norm folded into the convolution's passes (tc_conv.cu)
// ------------------------------------------------------------------------------------------
extern "C" int pdeqx_conv_gn_fusable(const pdeqx_conv_desc* d) {
  ConvGeom g;
  if (conv_geometry(d, &g) != PDEQX_OK) return 0;
  return tc_conv_gn_fusable(g) ? 1 : 0;
}

extern "C" int pdeqx_conv_gn_slots(const pdeqx_conv_desc* d) {
  ConvGeom g;
  if (conv_geometry(d, &g) != PDEQX_OK || !tc_conv_gn_fusable(g)) return 0;
  return tc_conv_gn_slots(g);
}

static int gn_geometry(const pdeqx_conv_desc* d, ConvGeom* g) {
  PDEQX_TRY(conv_geometry(d, g));
  if (!tc_conv_gn_fusable(*g)) {
    set_error("this convolution does not take the fused GroupNorm passes (pdeqx_conv_gn_fusable)");
    return PDEQX_ERR_UNSUPPORTED;
  }
  return PDEQX_OK;
}

extern "C" int pdeqx_conv_gn_fwd(const pdeqx_conv_desc* d, const float* x, const float* w, const float* b, const float* gn_stats,
                                 const float* gn_weight, const float* gn_bias, int32_t activation, float* y, float* out_parts,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(x && w && y && gn_stats, "null argument");
  PDEQX_CHECK_ARG(activation >= 0 && activation <= 2, "unknown activation %d", activation);
  ConvGeom g;
  PDEQX_TRY(gn_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  return tc_conv_gn_fwd(g, x, w, b, gn_stats, gn_weight, gn_bias, activation, y, out_parts, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

extern "C" int pdeqx_conv_gn_bwd_filter(const pdeqx_conv_desc* d, const float* x, const float* gy, const float* gn_stats,
                                        const float* gn_weight, const float* gn_bias, float* gw, float* gb, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(x && gy && gw && gn_stats, "null argument");
  ConvGeom g;
  PDEQX_TRY(gn_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  return tc_conv_gn_bwd_filter(g, x, gy, gn_stats, gn_weight, gn_bias, gw, gb, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int pdeqx_conv_gn_bwd_data(const pdeqx_conv_desc* d, const float* gy, const float* w, const float* gn_weight,
                                      const float* h, float* v, float* parts, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  PDEQX_CHECK_ARG(gy && w && h && v && parts, "null argument");
  ConvGeom g;
  PDEQX_TRY(gn_geometry(d, &g));
  if (g.batch == 0) return PDEQX_OK;
  return tc_conv_gn_bwd_data(g, gy, w, gn_weight, h, v, parts, workspace, workspace_bytes, (cudaStream_t)stream);
}
