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
#include "modemap.cuh"

namespace pdeqx {

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
  const bool valid = mode_to_weight(mm, p, g, loc);
  float2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  const int64_t wbase = (int64_t)g * O * I * mm.mprod + loc;
  for (int v = 0; v < (valid ? V : 0); ++v) {  // padding positions: zeros out
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
  if (!mode_to_weight(mm, p, g, loc)) return;  // padding position: no weight slot
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


// ---- tiled path: per-mode complex GEMM  C[m,n,p] = sum_k A(m,k,p) * B(n,k,p)^(conj) -------------------------------
// One CTA owns PT consecutive retained modes (one run along the contiguous mode axis, so the corner g and the weight
// offset are affine in the mode), 64 rows m and NT columns n; operands are staged through shared memory in K chunks,
// each thread keeps an 8 x 4 complex accumulator tile for one mode. The three uses differ only in where the operands
// live: mode-space arrays are complex-interleaved [row, J], weights are planar (re / im arrays) in the reference
// layout (G, o, i, *m).
//   mix    : M = b, N = u = o, K = v = i, B = w[o=n, i=k]            (planar), C interleaved
//   mix^H  : M = b, N = u = i, K = v = o, B = conj(w[o=k, i=n])      (planar), C interleaved
//   dW     : M = o, N = i,     K = b,     B = conj(xhat[b=k, i=n])   (interleaved), C planar
struct TiledArgs {
  const float2* a;   // interleaved
  int64_t a_sm, a_sk;  // strides (in float2) over m and k
  const float2* b2;  // interleaved B (dW)
  const float* br;   // planar B (mix)
  const float* bi;
  int64_t b_sn, b_sk;  // strides (float2 for interleaved / floats for planar) over n and k
  float2* c2;        // interleaved C (mix)
  float* cr;         // planar C (dW)
  float* ci;
  int64_t c_sm, c_sn;
  int M, N, K;
  int64_t J;
  int64_t w_corner;  // O * I * mprod: weight floats per corner
  ModeMap mm;
};

constexpr int TM_KC = 8;  // K chunk; a CTA covers 8 * RT rows (RT = rows per thread: small batches shrink the tile)

// cp.async (LDGSTS) with zero fill: `bytes` of 4 or 8 are copied when `pred`, else the destination is zero-filled
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int src_bytes = pred ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(gsrc), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_zfill16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int src_bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NS = cp.async stages: small batches (RT <= 2: a few samples per GPU) have almost no math per K chunk, so the kernel is a
// stream over the weights and needs more chunks in flight than the two that suffice when the FFMA work covers the loads
template <int PT, int RT, bool B_PLANAR, bool C_PLANAR, bool CONJ_B, int NS = (RT == 1 ? 4 : (RT == 2 ? 3 : 2))>
__global__ void __launch_bounds__(256, 2)
mode_gemm_tiled_kernel(TiledArgs t) {
  pdl_enter();
  constexpr int NQ = 32 / PT;   // n-quads per warp-row group ... threads = PT * 8 * NQ
  constexpr int NT = 4 * NQ;    // columns per CTA
  constexpr int TM_MT = 8 * RT; // rows per CTA
  constexpr int A_ELEMS = TM_KC * TM_MT * PT, B_ELEMS = TM_KC * NT * PT;
  // two stages of [A chunk: KC x MT x PT][B chunk: KC x NT x PT] float2, filled by cp.async one K chunk ahead of the math
  extern __shared__ float2 tile_smem[];
  const int tid = threadIdx.x;
  const int pm = tid % PT, mq = (tid / PT) % 8, nq = tid / (PT * 8);
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int m0 = blockIdx.y * TM_MT, n0 = blockIdx.z * NT;
  int g;
  int64_t loc;
  const bool valid = mode_to_weight(t.mm, p0, g, loc);  // a tile is all data or all padding (PT divides both mode counts)
  const int64_t woff = (int64_t)g * t.w_corner + loc;  // planar operands: offset of mode p0 inside an (o,i) plane
  float2 acc[RT][4];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = make_float2(0.f, 0.f);

  auto issue = [&](int k0, int stage) {
    float2* As = tile_smem + (size_t)stage * (A_ELEMS + B_ELEMS);
    float2* Bs = As + A_ELEMS;
    // A chunk: [KC][MT][PT] float2, PT contiguous in global memory
    for (int idx = tid; idx < A_ELEMS; idx += 256) {
      const int q = idx % PT, m = (idx / PT) % TM_MT, k = idx / (PT * TM_MT);
      const bool ok = m0 + m < t.M && k0 + k < t.K;
      const float2* src = ok ? t.a + (int64_t)(m0 + m) * t.a_sm + (int64_t)(k0 + k) * t.a_sk + p0 + q : t.a;
      cp_async_zfill<8>(As + idx, src, ok);
    }
    for (int idx = tid; idx < B_ELEMS; idx += 256) {
      const int q = idx % PT, n = (idx / PT) % NT, k = idx / (PT * NT);
      const bool ok = n0 + n < t.N && k0 + k < t.K;
      if (B_PLANAR) {
        break;  // (staged below as two planes with 16-byte copies)
      } else {
        const float2* src = ok ? t.b2 + (int64_t)(n0 + n) * t.b_sn + (int64_t)(k0 + k) * t.b_sk + p0 + q : t.b2;
        cp_async_zfill<8>(Bs + idx, src, ok);
      }
    }
    if (B_PLANAR) {
      // weights: re / im planes of the reference layout, PT consecutive modes = one contiguous run per (n, k): kept planar in
      // shared memory ([KC][NT][PT] floats each) so that a run moves as 16-byte copies -- at a few samples per GPU this kernel
      // is a pure stream over the weights and was bound by the number of 4-byte copy instructions
      float* Bre = (float*)Bs;
      float* Bim = Bre + B_ELEMS;
      constexpr int P4 = PT / 4;
      for (int idx = tid; idx < B_ELEMS / 4; idx += 256) {
        const int q4 = idx % P4, n = (idx / P4) % NT, k = idx / (P4 * NT);
        const bool ok = n0 + n < t.N && k0 + k < t.K;
        const int64_t off = ok ? (int64_t)(n0 + n) * t.b_sn + (int64_t)(k0 + k) * t.b_sk + woff + q4 * 4 : 0;
        const int dst = (k * NT + n) * PT + q4 * 4;
        cp_async_zfill16(Bre + dst, t.br + off, ok);
        cp_async_zfill16(Bim + dst, t.bi + off, ok);
      }
    }
    cp_async_commit();
  };

  const int nchunks = valid ? (t.K + TM_KC - 1) / TM_KC : 0;  // padding tile: interleaved outputs get zeros, planar ones nothing
  // NS - 1 chunks in flight ahead of the math; one commit group per iteration (possibly empty) keeps the wait count constant
  for (int c = 0; c < NS - 1; ++c) {
    if (c < nchunks) issue(c * TM_KC, c);
    else cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    if (c + NS - 1 < nchunks) issue((c + NS - 1) * TM_KC, (c + NS - 1) % NS);
    else cp_async_commit();
    cp_async_wait<NS - 1>();
    __syncthreads();
    const float2* As = tile_smem + (size_t)(c % NS) * (A_ELEMS + B_ELEMS);
    const float2* Bs = As + A_ELEMS;
#pragma unroll
    for (int k = 0; k < TM_KC; ++k) {
      float2 a[RT], b[4];
#pragma unroll
      for (int r = 0; r < RT; ++r) a[r] = As[(k * TM_MT + mq + 8 * r) * PT + pm];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (B_PLANAR) {
          const float* Bre = (const float*)Bs;
          b[cc] = make_float2(Bre[(k * NT + nq * 4 + cc) * PT + pm], Bre[B_ELEMS + (k * NT + nq * 4 + cc) * PT + pm]);
        } else {
          b[cc] = Bs[(k * NT + nq * 4 + cc) * PT + pm];
        }
        if (CONJ_B) b[cc].y = -b[cc].y;
      }
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          acc[r][cc].x = fmaf(a[r].x, b[cc].x, fmaf(-a[r].y, b[cc].y, acc[r][cc].x));
          acc[r][cc].y = fmaf(a[r].x, b[cc].y, fmaf(a[r].y, b[cc].x, acc[r][cc].y));
        }
    }
    __syncthreads();  // the stage is refilled by the prefetch of the next iteration
  }
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int m = m0 + mq + 8 * r;
    if (m >= t.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = n0 + nq * 4 + c;
      if (n >= t.N) continue;
      if (C_PLANAR) {
        if (!valid) continue;
        const int64_t off = (int64_t)m * t.c_sm + (int64_t)n * t.c_sn + woff + pm;
        t.cr[off] = acc[r][c].x;
        t.ci[off] = acc[r][c].y;
      } else {
        t.c2[(int64_t)m * t.c_sm + (int64_t)n * t.c_sn + p0 + pm] = acc[r][c];
      }
    }
  }
}

template <int PT, bool B_PLANAR, bool C_PLANAR, bool CONJ_B>
static bool launch_tiled_pt(const TiledArgs& t, cudaStream_t s) {
  // rows per thread: the smallest power of two whose 8 * RT-row tile covers M (at most 64 rows per CTA)
  int rt = 1;
  while (rt < 8 && 8 * rt < t.M) rt *= 2;
  const int nt = 4 * (32 / PT);
  dim3 grid((unsigned)(t.J / PT), (unsigned)ceil_div(t.M, 8 * rt), (unsigned)ceil_div(t.N, nt));
  if (grid.y > 65535 || grid.z > 65535) return false;
#define PDEQX_MODE_GEMM(RT_)                                                                                            \
  do {                                                                                                                 \
    const size_t smem = (size_t)(RT_ == 1 ? 4 : (RT_ == 2 ? 3 : 2)) * TM_KC * (8 * RT_ + nt) * PT * sizeof(float2);    \
    static DeviceOnce attr;                                                                                                \
    if (attr.first_use()) {                                                                                                       \
      cudaFuncSetAttribute(mode_gemm_tiled_kernel<PT, RT_, B_PLANAR, C_PLANAR, CONJ_B>,                                \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                                    \
                                                                                                                       \
    }                                                                                                                  \
    if (launch_pdl(mode_gemm_tiled_kernel<PT, RT_, B_PLANAR, C_PLANAR, CONJ_B>, grid, dim3(256), smem, s, t) != cudaSuccess) return false;                            \
  } while (0)
  switch (rt) {
    case 1: PDEQX_MODE_GEMM(1); break;
    case 2: PDEQX_MODE_GEMM(2); break;
    case 4: PDEQX_MODE_GEMM(4); break;
    default: PDEQX_MODE_GEMM(8); break;
  }
#undef PDEQX_MODE_GEMM
  return true;
}

template <bool B_PLANAR, bool C_PLANAR, bool CONJ_B>
static bool launch_tiled(const TiledArgs& t, cudaStream_t s) {
  const int mD = t.mm.m[t.mm.D - 1], nD = t.mm.nd[t.mm.D - 1];
  // few rows (a small per-GPU batch): the kernel streams the weights, so 16 consecutive modes per tile (64-byte runs of a
  // weight plane instead of 32-byte ones) matter more than the column reuse of the wider tile
  if (mD % 16 == 0 && nD % 16 == 0 && t.M <= 16) return launch_tiled_pt<16, B_PLANAR, C_PLANAR, CONJ_B>(t, s);
  if (mD % 8 == 0 && nD % 8 == 0) return launch_tiled_pt<8, B_PLANAR, C_PLANAR, CONJ_B>(t, s);
  if (mD % 4 == 0 && nD % 4 == 0) return launch_tiled_pt<4, B_PLANAR, C_PLANAR, CONJ_B>(t, s);
  return false;
}

int mode_mix(const pdeqx_spectral_plan* p, const float* in, const float* w_re, const float* w_im,
             bool transposed_conj, float* out, cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  const int O = d.c_out, I = d.c_in;
  const int U = transposed_conj ? I : O, V = transposed_conj ? O : I;
  if (tc_mode_gemm_available(p)) return tc_mode_mix(p, in, w_re, w_im, transposed_conj, out, s);
  {
    TiledArgs t{};
    t.mm = make_map(p);
    t.a = (const float2*)in; t.a_sm = (int64_t)V * p->J; t.a_sk = p->J;
    t.br = w_re; t.bi = w_im;
    // W(u,v): plain -> w[o=u, i=v]; transposed_conj -> conj(w[o=v, i=u])
    t.b_sn = transposed_conj ? t.mm.mprod : (int64_t)I * t.mm.mprod;
    t.b_sk = transposed_conj ? (int64_t)I * t.mm.mprod : t.mm.mprod;
    t.c2 = (float2*)out; t.c_sm = (int64_t)U * p->J; t.c_sn = p->J;
    t.M = d.batch; t.N = U; t.K = V; t.J = p->J; t.w_corner = (int64_t)O * I * t.mm.mprod;
    const bool aligned = (((uintptr_t)w_re | (uintptr_t)w_im) & 15) == 0 && t.mm.mprod % 4 == 0;  // 16-byte weight copies
    const bool ok = aligned && (transposed_conj ? launch_tiled<true, false, true>(t, s) : launch_tiled<true, false, false>(t, s));
    if (ok) { PDEQX_LAUNCHED(); return PDEQX_OK; }
  }
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
  if (tc_mode_gemm_available(p)) return tc_mode_mix_dw(p, ghat, xhat, gw_re, gw_im, s);
  {
    TiledArgs t{};
    t.mm = make_map(p);
    t.a = (const float2*)ghat; t.a_sm = p->J; t.a_sk = (int64_t)d.c_out * p->J;
    t.b2 = (const float2*)xhat; t.b_sn = p->J; t.b_sk = (int64_t)d.c_in * p->J;
    t.cr = gw_re; t.ci = gw_im; t.c_sm = (int64_t)d.c_in * t.mm.mprod; t.c_sn = t.mm.mprod;
    t.M = d.c_out; t.N = d.c_in; t.K = d.batch; t.J = p->J; t.w_corner = (int64_t)d.c_out * d.c_in * t.mm.mprod;
    if (launch_tiled<false, true, true>(t, s)) { PDEQX_LAUNCHED(); return PDEQX_OK; }
  }
  dim3 grid((unsigned)ceil_div(p->J, 16), (unsigned)ceil_div(d.c_out, 16), (unsigned)ceil_div(d.c_in, 16));
  PDEQX_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "mode_mix_dw: grid too large");
  mode_mix_dw_kernel<<<grid, 256, 0, s>>>((const float2*)ghat, (const float2*)xhat, gw_re, gw_im, d.batch, p->J,
                                           d.c_out, d.c_in, make_map(p));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx
