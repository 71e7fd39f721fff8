// tcgen05 (TF32) stages of the spectral block -- the three kernels that touch the full-size tensors.
//
//   tc_r2c   : X1[row, 0:2m]  = sum_w x[row, w] T[w, :]                 (truncated r2c DFT, first stage)
//   tc_slab  : y[b,o,l,w]     = epi( sum_k Z[b,o,l,k] Tc[k,w] + sum_i W[o,i] x[b,i,l,w] + bias[o] )
//              (inverse DFT along the contiguous axis + 1x1 bypass + bias + activation, one pass:
//               x is read once and y written once -- SURVEY.md App. D.1)
//   tc_wgrad : gW[o,i] = sum_{b,p} g[b,o,p] x[b,i,p],  gb[o] = sum_{b,p} g[b,o,p]   (bypass weight gradient)
//
// All three are persistent, warp-specialised kernels: warp 0 = TMA producer, warp 1 = tcgen05.mma
// issuer (one elected lane), warp 2 = TMEM allocator, warps 4.. = epilogue (tcgen05.ld -> registers ->
// coalesced global stores). Operand tiles are staged by TMA with the 128-byte swizzle and consumed
// through shared-memory descriptors; accumulators live in TMEM (double-buffered where the kernel has
// more than one output tile per CTA). They are HBM-bound by construction (16-36 flop/B, SURVEY.md
// App. B): the tensor pipe only has to stay under the memory time, which CUDA-core FFMA cannot.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "spectral_plan.h"
#include <type_traits>

#include "tc_common.cuh"

namespace pdeqx {
namespace tc {

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return PDEQX_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == TM_SW128_ATOM32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                             : (swizzle == TM_NONE ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu/%llu/%llu, box %u/%u/%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return PDEQX_ERR_CUDA;
  }
  return PDEQX_OK;
}

static float host_round_tf32(float v) {  // round-to-nearest (ties away), like cvt.rna.tf32.f32
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return v;
  u += 0x1000u;
  u &= 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

constexpr int kThreadsSlab = 640;   // warps 0-2 roles, warp 3 idle, warps 4-19 epilogue (4 per TMEM lane quarter)
constexpr int kThreadsR2c = 256;    // warps 0-2 roles, warp 3 idle, warps 4-7 epilogue
constexpr int kTmemCols = 256;

// ------------------------------------------------------------------------------------------
// epilogue math
// ------------------------------------------------------------------------------------------
template <bool FAST>
__device__ __forceinline__ float gelu_f(float x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float u = x * fmaf(x2, k01, k0);
  const float t = FAST ? tanh_fast(u) : tanh_accurate(u);
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
template <bool FAST>
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float u = x * fmaf(x2, k01, k0);
  const float t = FAST ? tanh_fast(u) : tanh_accurate(u);
  const float du = fmaf(x2, 3.0f * k01, k0);
  const float hx = 0.5f * x;
  // 0.5 (1 + t) + 0.5 x (1 - t^2) du
  return fmaf(hx * du, fmaf(-t, t, 1.0f), fmaf(0.5f, t, 0.5f));
}

// gelu(x) and gelu'(x) from one tanh
__device__ __forceinline__ void gelu_both_f(float x, float& y, float& dy) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float u = x * fmaf(x2, k01, k0);
  const float t = tanh_fast(u);
  const float du = fmaf(x2, 3.0f * k01, k0);
  const float hx = 0.5f * x;
  y = fmaf(hx, t, hx);
  dy = fmaf(hx * du, fmaf(-t, t, 1.0f), fmaf(0.5f, t, 0.5f));
}

// gelu / gelu' of two values at once on packed fp32 (same formulas, same per-lane rounding as gelu_f / gelu_grad_f<true>)
__device__ __forceinline__ f32x2 gelu2(f32x2 x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const f32x2 x2 = mul2(x, x);
  const f32x2 u = mul2(x, fma2(x2, pk2(k01, k01), pk2(k0, k0)));
  float u0, u1;
  upk2(u, u0, u1);
  const f32x2 t = pk2(tanh_fast(u0), tanh_fast(u1));
  const f32x2 hx = mul2(x, pk2(0.5f, 0.5f));
  return fma2(hx, t, hx);
}
__device__ __forceinline__ f32x2 gelu_grad2(f32x2 x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const f32x2 x2 = mul2(x, x);
  const f32x2 u = mul2(x, fma2(x2, pk2(k01, k01), pk2(k0, k0)));
  float u0, u1;
  upk2(u, u0, u1);
  const f32x2 t = pk2(tanh_fast(u0), tanh_fast(u1));
  const f32x2 du = fma2(x2, pk2(3.0f * k01, 3.0f * k01), pk2(k0, k0));
  const f32x2 hx = mul2(x, pk2(0.5f, 0.5f));
  const f32x2 omt2 = fma2(mul2(t, t), pk2(-1.0f, -1.0f), pk2(1.0f, 1.0f));  // 1 - t^2
  const f32x2 half = pk2(0.5f, 0.5f);
  return fma2(mul2(hx, du), omt2, fma2(half, t, half));  // 0.5 (1 + t) + 0.5 x (1 - t^2) du
}

// ==========================================================================================
// slab kernel
// ==========================================================================================
struct SlabParams {
  int lead, W, wt_per_row, c_in, c_out;
  int parts;  // epilogue warps per TMEM lane quarter: each owns c_out / parts channels (8, 16 or 32)
  int rpu;  // image rows per 128-pixel unit: 1 for W >= 128 (W / 128 units per row), 128 / W for narrower grids
  int has_bypass, mode, act, round_out;
  int stages;
  int64_t units;  // B * lead * wt_per_row
  const float* bias;
  const float* aux;  // mode 1: cotangent gy [B, c_out, lead, W]; mode 2: its single-channel factor gy1 [B, lead, W]
  float* red_w;        // mode 3: out[o] += sum_{b,pixel} result[b,o,pixel] * aux[b,pixel]   (aux = single-channel field)
  float* red_b;        // mode 3: out[o] += sum_{b,pixel} result[b,o,pixel]                 (may be null)
  const float* byp_u;  // BYP1: the bypass term is rank one, byp_a[o] * byp_u[b, pixel] (the block's input is a pointwise 1 -> C
  const float* byp_a;  //       lifting of the single-channel field byp_u; its constant part is folded into `bias`)
  const float* aux_w;  // mode 2: the per-channel factor, gy[b,o,p] = aux_w[o] * gy1[b,p] (cotangent behind a C -> 1 projection)
  float* y;
};

// smem carve-up (all tile regions 1024-B aligned)
struct SlabSmem {
  uint32_t tab_off, w_off, stage_off, stage_bytes, x_box_bytes, z_off_in_stage, epi_off, epi_buf_bytes, epi_bufs, bar_off, total;
  uint32_t chunk_bytes;
};

static SlabSmem slab_smem(int W, int c_in, int c_out, bool bypass, int stages, int epi_bufs = 2) {
  SlabSmem s{};
  uint32_t off = 0;
  const uint32_t rpu = W >= 128 ? 1 : 128 / W;
  s.tab_off = off;
  // W >= 128: one [128 w x 32 k] table tile (the grid is a multiple of W/128, so a CTA only ever sees one w-tile);
  // W < 128: the unit spans rpu rows and the table is block diagonal, one [128 pixels x 32 k] tile per row of the unit
  off += rpu * 128 * 128;
  s.w_off = off;
  off += bypass ? (uint32_t)((c_in + 31) / 32) * c_out * 128 : 0;
  s.x_box_bytes = (uint32_t)c_in * 128;
  s.z_off_in_stage = bypass ? 4 * s.x_box_bytes : 0;
  s.stage_bytes = s.z_off_in_stage + rpu * (uint32_t)c_out * 128;
  s.stage_bytes = (s.stage_bytes + 1023) & ~1023u;
  s.stage_off = off;
  off += s.stage_bytes * stages;
  // epilogue staging: `epi_bufs` output tiles [c_out x 128 w] fp32 (TMA store source / aux target), each stored as four
  // SWIZZLE_128B chunks [c_out rows x 32 w]: every epilogue warp owns one 128-byte-row box per part (its own TMA box)
  s.epi_off = off;
  s.chunk_bytes = (uint32_t)c_out * 128;
  s.epi_buf_bytes = 4 * s.chunk_bytes;
  off += epi_bufs > 0 ? (uint32_t)epi_bufs * s.epi_buf_bytes : 4096u;  // (modes 3 / 4 store no tile: 4 KB for the final reduction)
  s.epi_bufs = (uint32_t)epi_bufs;
  s.bar_off = off;
  off += 3072;  // barriers, tmem pointer, bias, rank-1 cotangent factor, rank-1 bypass factor
  s.total = off + 1024;  // slack for manual 1024-B alignment of the dynamic base
  return s;
}

template <int MODE, int ACT, bool BYP1>
__global__ void __launch_bounds__(kThreadsSlab, 1)
slab_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_z,
            const __grid_constant__ CUtensorMap tm_tab, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_aux, SlabParams p, SlabSmem L) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + L.bar_off);
  uint64_t* full = bars;                   // [stages]
  uint64_t* empty = bars + 8;              // [stages]
  uint64_t* tfull = bars + 16;             // [2]
  uint64_t* tempty = bars + 18;            // [2]
  uint64_t* tabfull = bars + 20;           // [1]
  uint32_t* tmem_ptr = (uint32_t*)(bars + 22);
  uint64_t* auxfull = bars + 24;           // [16 epilogue warps][3 buffers] cotangent chunk landed (MODE 1)
  float* bias_s = (float*)(bars + 72);     // [c_out] (<= 128 floats)
  float* auxw_s = bias_s + 128;            // [c_out] per-channel factor of a rank-1 cotangent (MODE 2)
  float* bypa_s = auxw_s + 128;            // [c_out] per-channel factor of a rank-1 bypass (BYP1)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_z);
    tma_prefetch_desc(&tm_tab);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4 * p.parts);
    mbar_init(&tempty[1], 4 * p.parts);
    mbar_init(tabfull, 1);
    for (int i = 0; i < 48; ++i) mbar_init(&auxfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
  if (warp == 3) {
    for (int i = lane; i < p.c_out; i += 32) {
      bias_s[i] = p.bias ? p.bias[i] : 0.f;
      auxw_s[i] = ((MODE == 2 || MODE == 4 || MODE == 5) && p.aux_w) ? p.aux_w[i] : 0.f;
      bypa_s[i] = (BYP1 && p.byp_a) ? p.byp_a[i] : 0.f;
    }
    if (lane == 0) {
      tma_prefetch_desc(&tm_y);
      tma_prefetch_desc(&tm_aux);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int64_t first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // constant operands: DFT table tiles and the bypass weight chunks
      const int wchunks = p.has_bypass ? (p.c_in + 31) / 32 : 0;
      mbar_expect_tx(tabfull, (uint32_t)p.rpu * 128 * 128 + (uint32_t)wchunks * p.c_out * 128);
      for (int c = 0; c < p.rpu; ++c)
        tma_load_2d(smem + L.tab_off + c * 128 * 128, &tm_tab, tabfull, 0, ((int)(first % p.wt_per_row) + c) * 128);
      for (int c = 0; c < wchunks; ++c) tma_load_2d(smem + L.w_off + c * p.c_out * 128, &tm_w, tabfull, c * 32, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t u = first; u < p.units; u += stride) {
        const int wt = (int)(u % p.wt_per_row);
        const int64_t bl = u / p.wt_per_row;
        const int lu = p.lead / p.rpu;
        const int l = (int)(bl % lu) * p.rpu;  // first image row of the unit
        const int b = (int)(bl / lu);
        mbar_wait_backoff(&empty[s], ph ^ 1);
        uint8_t* st = smem + L.stage_off + (size_t)s * L.stage_bytes;
        mbar_expect_tx(&full[s], (p.has_bypass ? 4 * L.x_box_bytes : 0) + (uint32_t)p.rpu * p.c_out * 128);
        if (p.has_bypass)
          for (int j = 0; j < 4; ++j) {
            const int px = wt * 128 + j * 32;  // pixel offset inside the unit's row group
            tma_load_3d(st + j * L.x_box_bytes, &tm_x, &full[s], px % p.W, l + px / p.W, b * p.c_in);
          }
        for (int c = 0; c < p.rpu; ++c)
          tma_load_3d(st + L.z_off_in_stage + c * p.c_out * 128, &tm_z, &full[s], 0, l + c, b * p.c_out);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_kk = idesc_tf32(128, p.c_out, false, false);
      const uint32_t idesc_mk = idesc_tf32(128, p.c_out, true, false);
      mbar_wait(tabfull, 0);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      const uint64_t kk0 = smem_desc_sw128(0, 0, 1024);                       // K-major operands (address field added below)
      const uint64_t mk0 = smem_desc_sw128_atom32(0, L.x_box_bytes, 512);     // MN-major x slab
      const uint32_t kk_hi = (uint32_t)(kk0 >> 32), kk_lo = (uint32_t)kk0, mk_hi = (uint32_t)(mk0 >> 32), mk_lo = (uint32_t)mk0;
      const uint32_t tab_lo = kk_lo + (smem_u32(smem + L.tab_off) >> 4), w_lo = kk_lo + (smem_u32(smem + L.w_off) >> 4);
      const uint32_t st16_0 = smem_u32(smem + L.stage_off) >> 4, stage16 = L.stage_bytes >> 4, z16 = L.z_off_in_stage >> 4;
      const uint32_t cbox16 = (uint32_t)p.c_out * 128 >> 4;
      for (int64_t u = first; u < p.units; u += stride) {
        mbar_wait(&tempty[a], aph ^ 1);
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(a * p.c_out);
        // descriptors: constant high words, low word = 14-bit start address (>> 4); per MMA only integer adds
        const uint32_t st16 = st16_0 + (uint32_t)s * stage16;
        // spectral part: A = table tile [128 w x 32 k] (K-major), B = Z box [c_out x 32 k] (K-major)
        for (int c = 0; c < p.rpu; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_tf32_lohi(d, tab_lo + (uint32_t)c * (128 * 128 >> 4) + k * 2, kk_hi,
                          kk_lo + st16 + z16 + (uint32_t)c * cbox16 + k * 2, kk_hi, idesc_kk, (c | k) ? 1u : 0u);
        // bypass part: A = x slab [128 w x c_in] (MN-major: w contiguous), B = W [c_out x c_in] (K-major)
        if (p.has_bypass) {
          const int ksteps = p.c_in / 8;
          for (int k = 0; k < ksteps; ++k)
            mma_tf32_lohi(d, mk_lo + st16 + (uint32_t)k * (1024 >> 4), mk_hi, w_lo + (uint32_t)(k >> 2) * cbox16 + (k & 3) * 2, kk_hi,
                          idesc_mk, 1u);
        }
        mma_commit(&empty[s]);
        mma_commit(&tfull[a]);
        if (++s == S) { s = 0; ph ^= 1; }
        if (++a == 2) { a = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4 && ((warp - 4) >> 2) < p.parts) {
    // ===================== epilogue: TMEM -> registers -> smem tile -> TMA store =====================
    // The kernel is bound by the latency of this stage (TMEM load -> activation math -> shared store), not by its issue
    // rate, so it gets 16 warps. Autonomous warps: warp (part, q) owns the [c_out/parts channels x 32 w] piece of the output tile that its TMEM
    // lane quarter produces -- one swizzled 128-byte-row box. It fills the piece, stores it with its own TMA store and
    // (mode 1) prefetches its own piece of the cotangent tile two units ahead: no CTA-level barrier in the loop.
    // The tile [c_out x 128 w] is laid out as four SWIZZLE_128B chunks [c_out rows x 32 w].
    const int q = warp & 3;            // TMEM lane quarter this warp may read = 32-w chunk of the tile
    const int part = (warp - 4) >> 2;  // which slice of the output channels
    const int ncol = p.c_out / p.parts;
    const int nc = ncol < 16 ? ncol : 16;  // valid columns of a 16-column TMEM load (8 when c_out = 32)
    const uint32_t bias_sa = smem_u32(bias_s), auxw_sa = smem_u32(auxw_s), bypa_sa = smem_u32(bypa_s);
    const int NB = (int)L.epi_bufs;  // 2 (MODE 0) or 3 (MODE 1: the cotangent piece is prefetched two units ahead)
    uint8_t* ebuf0 = smem + L.epi_off + (size_t)q * L.chunk_bytes + (size_t)part * ncol * 128;  // this warp's piece in buffer 0
    uint64_t* afull = auxfull + (part * 4 + q) * 3;
    int a = 0, e = 0;
    uint32_t aph = 0, eph = 0;
    // Unit coordinates advance incrementally (units stride by the grid size, a multiple of wt_per_row, so a CTA keeps its
    // w-tile and (row group, sample) move by a constant): this stage is bound by its instruction count and 64-bit
    // divisions per unit cost as much as the activation math.
    const int lu = p.lead / p.rpu;
    const int px = (int)(first % p.wt_per_row) * 128 + q * 32;  // this warp's pixels inside the unit's row group
    const int pxw = px % p.W, l_off = px / p.W;
    const int64_t dbl = stride / p.wt_per_row;
    const int adv_b = (int)(dbl / lu), adv_l = (int)(dbl % lu);
    struct Coord { int li, b; };
    auto advance = [&](Coord& c) {
      c.li += adv_l;
      c.b += adv_b;
      if (c.li >= lu) { c.li -= lu; ++c.b; }
    };
    Coord cur{(int)((first / p.wt_per_row) % lu), (int)((first / p.wt_per_row) / lu)};
    Coord nxt = cur;  // coordinates one unit ahead (rank-one field prefetch)
    advance(nxt);
    Coord nx2 = nxt;  // two units ahead (MODE 1: cotangent piece prefetch)
    advance(nx2);
    auto load_aux = [&](const Coord& c, int buf) {
      mbar_expect_tx(&afull[buf], (uint32_t)ncol * 128);
      tma_load_3d(ebuf0 + (size_t)buf * L.epi_buf_bytes, &tm_aux, &afull[buf], pxw, c.li * p.rpu + l_off, c.b * p.c_out + part * ncol);
    };
    float racc_w[2][16], racc_b[2][16];  // MODE 3 only (dead otherwise); c_out / parts <= 32 channels per warp
#pragma unroll
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int j = 0; j < 16; ++j) racc_w[ci][j] = racc_b[ci][j] = 0.f;
    if (MODE == 1 && lane == 0) {
      if (first < p.units) load_aux(cur, 0);
      if (first + stride < p.units) load_aux(nxt, 1);
    }
    // single-channel fields read per pixel (the rank-one factors): loaded ONE UNIT AHEAD -- this stage is latency bound and a
    // dependent global load at the top of every unit would sit on its critical path
    constexpr bool kUsesAux1 = MODE == 2 || MODE == 3 || MODE == 5;
    constexpr bool kRoundOut = (MODE == 0 && ACT != PDEQX_ACT_IDENTITY) || MODE == 1 || MODE == 2 || MODE == 5;
    auto field_index = [&](const Coord& c) -> int64_t {
      return ((int64_t)c.b * p.lead + c.li * p.rpu + l_off) * p.W + pxw + lane;
    };
    float sv_next = 0.f, su_next = 0.f;
    if (first < p.units) {
      const int64_t fi = field_index(cur);
      if (kUsesAux1) sv_next = __ldg(p.aux + fi);
      if (BYP1) su_next = __ldg(p.byp_u + fi);
    }
    for (int64_t u = first; u < p.units; u += stride, cur = nxt, nxt = nx2, advance(nx2)) {
      const int l = cur.li * p.rpu + l_off;
      const int b = cur.b;
      uint8_t* piece = ebuf0 + (size_t)e * L.epi_buf_bytes;
      const float sv = sv_next;  // MODE 2 / 3 / 5: this lane's pixel of the single-channel cotangent factor / field
      const float su = su_next;  // BYP1: this lane's pixel of the single-channel field behind the lifting
      if ((kUsesAux1 || BYP1) && u + stride < p.units) {
        const int64_t fi = field_index(nxt);
        if (kUsesAux1) sv_next = __ldg(p.aux + fi);
        if (BYP1) su_next = __ldg(p.byp_u + fi);
      }
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      if (MODE == 1) {
        mbar_wait(&afull[e], eph);
      } else {
        // this warp's store that last used the piece (NB units ago) must have finished reading it
        if (MODE != 3 && MODE != 4) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
      }
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.c_out + part * ncol);
      if (MODE == 3) {
        // nothing is stored: the result only feeds the gradients of the 1 -> C pointwise lifting that produced this
        // block's input from the single-channel field `aux` -- per-channel dot products kept in registers across units
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c0 = ci * 16;
          if (c0 < ncol) {
            float v[16];
            tmem_ld16(t0 + c0, v);
            tmem_ld_wait();
            if (c0 + 16 >= ncol) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[a]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j < nc) {
                const float r = v[j] + bias_s[part * ncol + c0 + j];
                racc_w[ci][j] = fmaf(r, sv, racc_w[ci][j]);
                racc_b[ci][j] += r;
              }
            }
          }
        }
        if (++a == 2) { a = 0; aph ^= 1; }
        continue;
      }
      // element (o_local, w = 32 q + lane): row o_local of the piece, 16-byte units XOR-swizzled by (row & 7)
      const uint32_t col = smem_u32(piece);
      float pacc = 0.f;  // MODE 4: projected value of this lane's pixel over this warp's channels
      const uint32_t lane_b = (uint32_t)lane * 4;
      // one 16-column TMEM load of which NC (8 or 16, compile time: the element loop is straight-line code, so the
      // independent activation chains of a chunk overlap) columns are this warp's channels
      auto chunk = [&](int c0, auto nc_tag) {
        constexpr int NC = decltype(nc_tag)::value;
        float v[16];
        tmem_ld16(t0 + c0, v);
        tmem_ld_wait();
        if (c0 + 16 >= ncol) {  // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[a]);
        }
        // per-channel constants of this chunk (shared-space vector loads)
        const uint32_t cidx = (uint32_t)(part * ncol + c0) * 4;
        float bias_c[NC], bypa_c[NC], auxw_c[NC];
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
          ld_shared_const_v4(bias_sa + cidx + j * 4, bias_c + j);
          if (BYP1) ld_shared_const_v4(bypa_sa + cidx + j * 4, bypa_c + j);
          if (MODE == 2 || MODE == 4 || MODE == 5) ld_shared_const_v4(auxw_sa + cidx + j * 4, auxw_c + j);
        }
        const uint32_t row0 = col + (uint32_t)c0 * 128;
        float g[NC];
        if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < NC; ++j) g[j] = ld_shared_f32(row0 + (uint32_t)j * 128 + (lane_b ^ ((uint32_t)(j & 7) << 4)));
        } else if (MODE == 2 || MODE == 5) {
#pragma unroll
          for (int j = 0; j < NC; ++j) g[j] = auxw_c[j] * sv;
        }
        if (ACT == PDEQX_ACT_GELU_TANH && (MODE == 0 || MODE == 4 || MODE == 1 || MODE == 2)) {
          // packed path: two channels per instruction (bias add, rank-one bypass, gelu / gelu')
          const f32x2 su2 = pk2(su, su);
#pragma unroll
          for (int j = 0; j < NC; j += 2) {
            f32x2 val = add2(pk2(v[j], v[j + 1]), pk2(bias_c[j], bias_c[j + 1]));
            if (BYP1) val = fma2(pk2(bypa_c[j], bypa_c[j + 1]), su2, val);
            float r0, r1;
            if (MODE == 4) {
              upk2(gelu2(val), r0, r1);
              pacc = fmaf(auxw_c[j], r0, pacc);
              pacc = fmaf(auxw_c[j + 1], r1, pacc);
              continue;
            }
            f32x2 out = MODE == 0 ? gelu2(val) : mul2(pk2(g[j], g[j + 1]), gelu_grad2(val));
            if (kRoundOut) out = round_tf32x2(out);
            upk2(out, r0, r1);
            st_shared_f32(row0 + (uint32_t)j * 128 + (lane_b ^ ((uint32_t)(j & 7) << 4)), r0);
            st_shared_f32(row0 + (uint32_t)(j + 1) * 128 + (lane_b ^ ((uint32_t)((j + 1) & 7) << 4)), r1);
          }
          return;
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          float val = v[j] + bias_c[j];
          if (BYP1) val = fmaf(bypa_c[j], su, val);
          float r;
          if (MODE == 0 || MODE == 4) {
            if (ACT == PDEQX_ACT_GELU_TANH) r = gelu_f<true>(val);
            else if (ACT == PDEQX_ACT_RELU) r = fmaxf(val, 0.f);
            else r = val;
            if (MODE == 4) {  // the activated value only feeds the C -> 1 projection: nothing is stored
              pacc = fmaf(auxw_c[j], r, pacc);
              continue;
            }
          } else if (MODE == 5) {
            // rank-one cotangent AND the projection's weight gradient sum_pixels act(pre)[o] * gy1 (the activated output was
            // never stored: it is recomputed here from the same pre-activation; ncol <= 16 in this mode)
            float yv, dy;
            if (ACT == PDEQX_ACT_GELU_TANH) gelu_both_f(val, yv, dy);
            else { yv = fmaxf(val, 0.f); dy = val > 0.f ? 1.f : 0.f; }
            r = g[j] * dy;
            racc_w[0][j] = fmaf(yv, sv, racc_w[0][j]);
          } else {
            if (ACT == PDEQX_ACT_GELU_TANH) r = g[j] * gelu_grad_f<true>(val);
            else if (ACT == PDEQX_ACT_RELU) r = val > 0.f ? g[j] : 0.f;
            else r = g[j];
          }
          // what this kernel stores is read next as a tcgen05 operand (the activated output by the next block's bypass and
          // first-DFT-stage MMAs, the cotangent by the weight-gradient and data-gradient MMAs), and the tensor core TRUNCATES
          // fp32 to tf32 (a bias of about -2^-12 per operand that adds up coherently through the layers): round to nearest here
          if (kRoundOut) r = round_tf32(r);
          st_shared_f32(row0 + (uint32_t)j * 128 + (lane_b ^ ((uint32_t)(j & 7) << 4)), r);
        }
      };
      if (nc == 16) {
        for (int c0 = 0; c0 < ncol; c0 += 16) {
          if (c0 + 16 <= ncol) chunk(c0, std::integral_constant<int, 16>{});
          else chunk(c0, std::integral_constant<int, 8>{});  // 24 channels per warp (c_out = 96): the tail chunk has 8
        }
      } else {
        chunk(0, std::integral_constant<int, 8>{});
      }
      if (MODE == 4) {  // this warp's channel slice of the projected pixel: the slices meet in the output by atomic adds
        atomicAdd(p.y + ((int64_t)b * p.lead + l) * p.W + pxw + lane, pacc);
        if (++a == 2) { a = 0; aph ^= 1; }
        continue;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tm_y, piece, pxw, l, b * p.c_out + part * ncol);
        tma_store_commit();
        if (MODE == 1) {
          const int64_t un = u + 2 * stride;
          if (un < p.units) {  // cotangent piece two units ahead, into the buffer whose store (previous unit) has drained
            tma_store_wait_read<1>();
            load_aux(nx2, e == 0 ? NB - 1 : e - 1);
          }
        }
      }
      if (++a == 2) { a = 0; aph ^= 1; }
      if (++e == NB) { e = 0; eph ^= 1; }
    }
    if (MODE == 3 || MODE == 5) {
      // warp-level reduction over the 32 pixels, then the four lane-quarter warps of a channel slice are combined in
      // shared memory (the output tiles are idle in this mode): ONE atomic per channel and CTA
      float* red_s = (float*)(smem + L.epi_off);  // [2 (w, b)][4 q][c_out]
      if (MODE == 5) {  // the tiles were in use here: every warp's stores must have read them before they are reused
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        named_bar_sync(1, 128 * p.parts);
      }
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        if (ci * 16 < ncol) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j >= nc) break;
            float vw = racc_w[ci][j], vb = racc_b[ci][j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              vw += __shfl_xor_sync(0xffffffffu, vw, off);
              vb += __shfl_xor_sync(0xffffffffu, vb, off);
            }
            if (lane == 0) {
              red_s[(0 * 4 + q) * p.c_out + part * ncol + ci * 16 + j] = vw;
              red_s[(1 * 4 + q) * p.c_out + part * ncol + ci * 16 + j] = vb;
            }
          }
        }
      }
      named_bar_sync(1, 128 * p.parts);
      if (q == 0) {
        for (int c = lane; c < ncol; c += 32) {
          const int ch = part * ncol + c;
          float vw = 0.f, vb = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            vw += red_s[(0 * 4 + qq) * p.c_out + ch];
            vb += red_s[(1 * 4 + qq) * p.c_out + ch];
          }
          atomicAdd(p.red_w + ch, vw);
          if (p.red_b) atomicAdd(p.red_b + ch, vb);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ==========================================================================================
// chain kernel: the data-gradient pass of block k FUSED with the pre-activation-cotangent pass of block k-1
//
//   gx_k[b,c,l,w]    = sum_kk Z'_k[b,c,l,kk] Tb[kk,w] + sum_o W_k[o,c] g_k[b,o,l,w]              (accumulator 1)
//   pre_{k-1}[b,c,l,w] = sum_kk Z_{k-1}[b,c,l,kk] Ti[kk,w] + sum_i W_{k-1}[c,i] x_{k-1}[b,i,l,w] + bias (accumulator 2)
//   g_{k-1}[b,c,l,w] = gx_k * act'(pre_{k-1})                                                     (epilogue, stored)
//
// gx_k (= the cotangent gy_{k-1} of the upstream block's output) is never written to HBM and never read back: per
// block boundary one full-size tensor write and one read disappear (2 of the 7 tensor passes a block's reverse pass made).
// Every operand layout is the slab kernel's: two load sub-stages per unit ([g_k slab | Z'_k] then [x_{k-1} slab | Z_{k-1}])
// walk through one ring; both accumulators of a unit sit side by side in TMEM (2 x C columns, double buffered).
// All four channel counts are equal (C <= 64). BYP1: the upstream block's input is the 1 -> C lifting of a single-channel
// field (rank-one bypass in the epilogue, no x slab in its sub-stage).
// ==========================================================================================
struct ChainParams {
  int lead, W, wt_per_row, C;
  int rpu, stages;
  int64_t units;
  const float* bias;   // upstream block: bypass bias (BYP1: the folded constant part)
  const float* byp_u;  // BYP1: single-channel field [B, lead, W]
  const float* byp_a;  // BYP1: per-channel factor
};

struct ChainSmem {
  uint32_t tab1_off, tab2_off, w1_off, w2_off, stage_off, stage_bytes, x_box_bytes, z_off_in_stage, epi_off, chunk_bytes, bar_off, total;
};

static ChainSmem chain_smem(int W, int C, int stages) {
  ChainSmem s{};
  uint32_t off = 0;
  const uint32_t rpu = W >= 128 ? 1 : 128 / W;
  s.tab1_off = off;
  off += rpu * 128 * 128;
  s.tab2_off = off;
  off += rpu * 128 * 128;
  const uint32_t wbytes = (uint32_t)((C + 31) / 32) * C * 128;
  s.w1_off = off;
  off += wbytes;
  s.w2_off = off;
  off += wbytes;
  s.x_box_bytes = (uint32_t)C * 128;
  s.z_off_in_stage = 4 * s.x_box_bytes;
  s.stage_bytes = (s.z_off_in_stage + rpu * (uint32_t)C * 128 + 1023) & ~1023u;
  s.stage_off = off;
  off += s.stage_bytes * stages;
  s.epi_off = off;  // ONE output tile [C x 128 w] as four SWIZZLE_128B chunks [C rows x 32 w]
  s.chunk_bytes = (uint32_t)C * 128;
  off += 4 * s.chunk_bytes;
  s.bar_off = off;
  off += 2048;
  s.total = off + 1024;
  return s;
}

template <int ACT, bool BYP1>
__global__ void __launch_bounds__(kThreadsSlab, 1)
chain_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z1,
             const __grid_constant__ CUtensorMap tm_x2, const __grid_constant__ CUtensorMap tm_z2,
             const __grid_constant__ CUtensorMap tm_tab1, const __grid_constant__ CUtensorMap tm_tab2,
             const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
             const __grid_constant__ CUtensorMap tm_y, ChainParams p, ChainSmem L) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + L.bar_off);
  uint64_t* full = bars;          // [stages]
  uint64_t* empty = bars + 8;     // [stages]
  uint64_t* tfull = bars + 16;    // [2]
  uint64_t* tempty = bars + 18;   // [2]
  uint64_t* tabfull = bars + 20;  // [1]
  uint32_t* tmem_ptr = (uint32_t*)(bars + 22);
  float* bias_s = (float*)(bars + 24);  // [C] (<= 64 floats)
  float* bypa_s = bias_s + 64;          // [C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int C = p.C;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_z1);
    tma_prefetch_desc(&tm_x2);
    tma_prefetch_desc(&tm_z2);
    tma_prefetch_desc(&tm_tab1);
    tma_prefetch_desc(&tm_tab2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 16);
    mbar_init(&tempty[1], 16);
    mbar_init(tabfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
  if (warp == 3) {
    for (int i = lane; i < C; i += 32) {
      bias_s[i] = p.bias ? p.bias[i] : 0.f;
      bypa_s[i] = (BYP1 && p.byp_a) ? p.byp_a[i] : 0.f;
    }
    if (lane == 0) {
      tma_prefetch_desc(&tm_w1);
      tma_prefetch_desc(&tm_w2);
      tma_prefetch_desc(&tm_y);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int wchunks = (C + 31) / 32;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(tabfull, 2u * (uint32_t)p.rpu * 128 * 128 + (BYP1 ? 1u : 2u) * (uint32_t)wchunks * C * 128);
      const int wt0 = (int)(first % p.wt_per_row);
      for (int c = 0; c < p.rpu; ++c) {
        tma_load_2d(smem + L.tab1_off + c * 128 * 128, &tm_tab1, tabfull, 0, (wt0 + c) * 128);
        tma_load_2d(smem + L.tab2_off + c * 128 * 128, &tm_tab2, tabfull, 0, (wt0 + c) * 128);
      }
      for (int c = 0; c < wchunks; ++c) {
        tma_load_2d(smem + L.w1_off + c * C * 128, &tm_w1, tabfull, c * 32, 0);
        if (!BYP1) tma_load_2d(smem + L.w2_off + c * C * 128, &tm_w2, tabfull, c * 32, 0);
      }
      int s = 0;
      uint32_t ph = 0;
      const int lu = p.lead / p.rpu;
      for (int64_t u = first; u < p.units; u += stride) {
        const int wt = (int)(u % p.wt_per_row);
        const int64_t bl = u / p.wt_per_row;
        const int l = (int)(bl % lu) * p.rpu;
        const int b = (int)(bl / lu);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const bool has_x = half == 0 || !BYP1;
          mbar_wait_backoff(&empty[s], ph ^ 1);
          uint8_t* st = smem + L.stage_off + (size_t)s * L.stage_bytes;
          mbar_expect_tx(&full[s], (has_x ? 4 * L.x_box_bytes : 0) + (uint32_t)p.rpu * C * 128);
          if (has_x)
            for (int j = 0; j < 4; ++j) {
              const int px = wt * 128 + j * 32;
              tma_load_3d(st + j * L.x_box_bytes, half == 0 ? &tm_g : &tm_x2, &full[s], px % p.W, l + px / p.W, b * C);
            }
          for (int c = 0; c < p.rpu; ++c)
            tma_load_3d(st + L.z_off_in_stage + c * C * 128, half == 0 ? &tm_z1 : &tm_z2, &full[s], 0, l + c, b * C);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_kk = idesc_tf32(128, C, false, false);
      const uint32_t idesc_mk = idesc_tf32(128, C, true, false);
      mbar_wait(tabfull, 0);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      const uint64_t kk0 = smem_desc_sw128(0, 0, 1024);
      const uint64_t mk0 = smem_desc_sw128_atom32(0, L.x_box_bytes, 512);
      const uint32_t kk_hi = (uint32_t)(kk0 >> 32), kk_lo = (uint32_t)kk0, mk_hi = (uint32_t)(mk0 >> 32), mk_lo = (uint32_t)mk0;
      const uint32_t tab_lo[2] = {kk_lo + (smem_u32(smem + L.tab1_off) >> 4), kk_lo + (smem_u32(smem + L.tab2_off) >> 4)};
      const uint32_t w_lo[2] = {kk_lo + (smem_u32(smem + L.w1_off) >> 4), kk_lo + (smem_u32(smem + L.w2_off) >> 4)};
      const uint32_t st16_0 = smem_u32(smem + L.stage_off) >> 4, stage16 = L.stage_bytes >> 4, z16 = L.z_off_in_stage >> 4;
      const uint32_t cbox16 = (uint32_t)C * 128 >> 4;
      const int ksteps = C / 8;
      for (int64_t u = first; u < p.units; u += stride) {
        mbar_wait(&tempty[a], aph ^ 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(a * 2 * C + half * C);
          const uint32_t st16 = st16_0 + (uint32_t)s * stage16;
          for (int c = 0; c < p.rpu; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_tf32_lohi(d, tab_lo[half] + (uint32_t)c * (128 * 128 >> 4) + k * 2, kk_hi,
                            kk_lo + st16 + z16 + (uint32_t)c * cbox16 + k * 2, kk_hi, idesc_kk, (c | k) ? 1u : 0u);
          if (half == 0 || !BYP1)
            for (int k = 0; k < ksteps; ++k)
              mma_tf32_lohi(d, mk_lo + st16 + (uint32_t)k * (1024 >> 4), mk_hi, w_lo[half] + (uint32_t)(k >> 2) * cbox16 + (k & 3) * 2,
                            kk_hi, idesc_mk, 1u);
          mma_commit(&empty[s]);
          if (++s == S) { s = 0; ph ^= 1; }
        }
        mma_commit(&tfull[a]);
        if (++a == 2) { a = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: two accumulators -> g = acc1 * act'(acc2 + bias) -> smem piece -> TMA store =====================
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    const int ncol = C / 4;  // 16 (C = 64) or 8 (C = 32)
    const uint32_t bias_sa = smem_u32(bias_s), bypa_sa = smem_u32(bypa_s);
    uint8_t* piece = smem + L.epi_off + (size_t)q * L.chunk_bytes + (size_t)part * ncol * 128;
    const uint32_t col = smem_u32(piece);
    int a = 0;
    uint32_t aph = 0;
    const int lu = p.lead / p.rpu;
    const int px = (int)(first % p.wt_per_row) * 128 + q * 32;
    const int pxw = px % p.W, l_off = px / p.W;
    const int64_t dbl = stride / p.wt_per_row;
    const int adv_b = (int)(dbl / lu), adv_l = (int)(dbl % lu);
    int cli = (int)((first / p.wt_per_row) % lu), cb = (int)((first / p.wt_per_row) / lu);
    int nli = cli + adv_l, nb = cb + adv_b;
    if (nli >= lu) { nli -= lu; ++nb; }
    float su_next = 0.f;
    if (BYP1 && first < p.units) su_next = __ldg(p.byp_u + ((int64_t)cb * p.lead + cli * p.rpu + l_off) * p.W + pxw + lane);
    const uint32_t lane_b = (uint32_t)lane * 4;
    for (int64_t u = first; u < p.units; u += stride) {
      const int l = cli * p.rpu + l_off;
      const int b = cb;
      const float su = su_next;
      if (BYP1 && u + stride < p.units) su_next = __ldg(p.byp_u + ((int64_t)nb * p.lead + nli * p.rpu + l_off) * p.W + pxw + lane);
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      if (lane == 0) tma_store_wait_read<0>();  // this warp's previous store has read the piece
      __syncwarp();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 2 * C + part * ncol);
      auto chunk = [&](auto nc_tag) {
        constexpr int NC = decltype(nc_tag)::value;
        float v1[16], v2[16];
        tmem_ld16(t0, v1);
        tmem_ld16(t0 + (uint32_t)C, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);
        const uint32_t cidx = (uint32_t)(part * ncol) * 4;
        float bias_c[NC], bypa_c[NC];
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
          ld_shared_const_v4(bias_sa + cidx + j * 4, bias_c + j);
          if (BYP1) ld_shared_const_v4(bypa_sa + cidx + j * 4, bypa_c + j);
        }
        if (ACT == PDEQX_ACT_GELU_TANH) {  // packed fp32: two channels per instruction
          const f32x2 su2 = pk2(su, su);
#pragma unroll
          for (int j = 0; j < NC; j += 2) {
            f32x2 pre = add2(pk2(v2[j], v2[j + 1]), pk2(bias_c[j], bias_c[j + 1]));
            if (BYP1) pre = fma2(pk2(bypa_c[j], bypa_c[j + 1]), su2, pre);
            float r0, r1;
            // read next as a tcgen05 operand (weight gradient, data gradient): round, do not let it truncate
            upk2(round_tf32x2(mul2(pk2(v1[j], v1[j + 1]), gelu_grad2(pre))), r0, r1);
            st_shared_f32(col + (uint32_t)j * 128 + (lane_b ^ ((uint32_t)(j & 7) << 4)), r0);
            st_shared_f32(col + (uint32_t)(j + 1) * 128 + (lane_b ^ ((uint32_t)((j + 1) & 7) << 4)), r1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            float pre = v2[j] + bias_c[j];
            if (BYP1) pre = fmaf(bypa_c[j], su, pre);
            const float r = round_tf32(pre > 0.f ? v1[j] : 0.f);
            st_shared_f32(col + (uint32_t)j * 128 + (lane_b ^ ((uint32_t)(j & 7) << 4)), r);
          }
        }
      };
      if (ncol == 16) chunk(std::integral_constant<int, 16>{});
      else chunk(std::integral_constant<int, 8>{});
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tm_y, piece, pxw, l, b * C + part * ncol);
        tma_store_commit();
      }
      if (++a == 2) { a = 0; aph ^= 1; }
      cli = nli;
      cb = nb;
      nli += adv_l;
      nb += adv_b;
      if (nli >= lu) { nli -= lu; ++nb; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ==========================================================================================
// r2c kernel: out[row, 0:n_valid] = sum_k x[row, k] T[k, :]
// ==========================================================================================
struct R2cParams {
  int64_t rows, tiles;
  int kchunks, n_valid, out_ld, stages, round_out;
  float* out;
};

__global__ void __launch_bounds__(kThreadsR2c, 1)
r2c_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_tab, R2cParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // [table: kchunks x 4 KB][stages x 16 KB][barriers]
  const uint32_t tab_bytes = (uint32_t)p.kchunks * 32 * 128;
  uint8_t* stage0 = smem + tab_bytes;
  uint64_t* bars = (uint64_t*)(stage0 + (size_t)p.stages * 128 * 128);
  uint64_t* full = bars;
  uint64_t* empty = bars + 16;
  uint64_t* tfull = bars + 32;
  uint64_t* tempty = bars + 34;
  uint64_t* tabfull = bars + 36;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 38);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_tab);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4);
    mbar_init(&tempty[1], 4);
    mbar_init(tabfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<64>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int64_t first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(tabfull, tab_bytes);
      for (int c = 0; c < p.kchunks; ++c) tma_load_2d(smem + c * 32 * 128, &tm_tab, tabfull, c * 32, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = first; t < p.tiles; t += stride) {
        for (int c = 0; c < p.kchunks; ++c) {
          mbar_wait_backoff(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], 128 * 128);
          tma_load_2d(stage0 + (size_t)s * 128 * 128, &tm_x, &full[s], c * 32, (int)(t * 128));
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, 32, false, false);
      mbar_wait(tabfull, 0);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      for (int64_t t = first; t < p.tiles; t += stride) {
        mbar_wait(&tempty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(a * 32);
        for (int c = 0; c < p.kchunks; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t xa = smem_u32(stage0 + (size_t)s * 128 * 128);
          const uint32_t tb = smem_u32(smem + c * 32 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_tf32(d, smem_desc_sw128(xa + k * 32, 0, 1024), smem_desc_sw128(tb + k * 32, 0, 1024), idesc, (c | k) != 0);
          mma_commit(&empty[s]);
          if (++s == S) { s = 0; ph ^= 1; }
        }
        mma_commit(&tfull[a]);
        if (++a == 2) { a = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int a = 0;
    uint32_t aph = 0;
    for (int64_t t = first; t < p.tiles; t += stride) {
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      const int64_t row = t * 128 + q * 32 + lane;
      if (row < p.rows) {
        float* o = p.out + row * p.out_ld;
        if (p.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
        }
        if (p.n_valid == 32 && (p.out_ld & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *(float4*)(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < p.n_valid) o[j] = v[j];
        }
      }
      if (++a == 2) { a = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

// ==========================================================================================
// bypass weight gradient: gW[o,i] += sum_p g[o,p] x[i,p], gb[o] += sum_p g[o,p]   (per CTA partial, atomics)
// ==========================================================================================
struct WgradParams {
  int c_out, c_in, chunks_per_row, lead, stages;
  int c_in_pad;  // rows of the x box (>= c_in, a multiple of 8): a single-channel x (the field in front of a 1 -> C lifting)
                 // is loaded as a 16-row box whose other rows (the next samples' fields) only feed accumulator columns
                 // that are never read back
  int64_t rows;  // batch * lead image rows; a unit = one row = chunks_per_row stages of 32 pixels
  float* gw;
  float* gb;
  // fused first stage of the adjoint-of-inverse DFT of g (optional): x1[(b * c_out + o) * lead + l][0:32] = sum_w g[b,o,l,w] T[w,:]
  float* x1;
  int round_x1;  // x1 feeds a tcgen05 c2c stage (D >= 2): round to tf32 here
  int m64;       // c_out <= 64: M = 64 MMAs (A = the g rows only: half the A operand bytes; accumulator row r sits in TMEM lane
                 // 32 * (r / 16) + r % 16, i.e. 16 rows per lane quarter)
};

constexpr int kWgGroup = 4;  // image rows whose first-DFT-stage results share one accumulator tile (one drain per group)

__global__ void __launch_bounds__(kThreadsR2c, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x,
             const __grid_constant__ CUtensorMap tm_tab, WgradParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // [r2c table: chunks_per_row x 4 KB (fused only)] [stages] ; stage = [g box: c_out x 128 B][x box: c_in x 128 B][ones: 16 x 128 B];
  // A = 128 rows from the g box on
  const bool fused = p.x1 != nullptr;
  const uint32_t tab_bytes = fused ? (uint32_t)p.chunks_per_row * 4096 : 0;
  uint8_t* stage0 = smem + tab_bytes;
  const uint32_t g_bytes = (uint32_t)p.c_out * 128, x_bytes = (uint32_t)p.c_in_pad * 128;
  const uint32_t stage_bytes = g_bytes + x_bytes + 2048;
  uint64_t* bars = (uint64_t*)(stage0 + (size_t)p.stages * stage_bytes + 16384);  // 16 KB slack: A tile of the last stage
  uint64_t* full = bars;
  uint64_t* empty = bars + 16;
  uint64_t* done = bars + 32;
  uint64_t* tabfull = bars + 33;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 34);
  uint64_t* x1full = bars + 36;   // [2]
  uint64_t* x1empty = bars + 38;  // [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_tab);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], p.x1 != nullptr ? 2 : 1);  // fused: two issuing warps read every stage
    }
    mbar_init(done, 1);
    mbar_init(tabfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x1full[i], 1);
      mbar_init(&x1empty[i], p.m64 ? 4 : 2);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);  // 128 (gW | gb) + 2 x kWgGroup x 32 (first-DFT-stage tiles)
  // ones rows (bias-gradient columns) and a defined value everywhere an A/B descriptor can reach
  if (p.m64) {
    // M = 64: A is exactly the g box and B the x box + the ones rows -- TMA rewrites both boxes before every use, so only the
    // ones rows need a value (the full fill below is ~200 KB of shared-memory stores per CTA: 3-4 us of a 60 us launch at 8
    // samples per GPU)
    for (uint32_t i = threadIdx.x; i < (uint32_t)S * 512; i += blockDim.x)
      ((float*)(stage0 + (size_t)(i >> 9) * stage_bytes + g_bytes + x_bytes))[i & 511] = 1.0f;
  } else {
    for (uint32_t i = threadIdx.x; i < ((uint32_t)S * stage_bytes + 16384) / 4; i += blockDim.x) {
      const uint32_t in_stage = (i * 4) % stage_bytes;
      ((float*)stage0)[i] = (i * 4 < (uint32_t)S * stage_bytes && in_stage >= g_bytes + x_bytes) ? 1.0f : 0.0f;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t groups = (p.rows + kWgGroup - 1) / kWgGroup;  // a CTA walks groups of kWgGroup consecutive image rows
  const int N = p.c_in_pad + 16;
  const int W = p.chunks_per_row * 32;

  if (warp == 0) {
    if (lane == 0) {
      if (fused) {
        mbar_expect_tx(tabfull, tab_bytes);
        for (int c = 0; c < p.chunks_per_row; ++c) tma_load_2d(smem + c * 4096, &tm_tab, tabfull, c * 32, 0);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int64_t gi = first; gi < groups; gi += stride) {
        for (int rr = 0; rr < kWgGroup; ++rr) {
          const int64_t r = gi * kWgGroup + rr;
          if (r >= p.rows) break;
          const int b = (int)(r / p.lead);
          const int l = (int)(r % p.lead);
          for (int c = 0; c < p.chunks_per_row; ++c) {
            mbar_wait_backoff(&empty[s], ph ^ 1);
            uint8_t* st = stage0 + (size_t)s * stage_bytes;
            mbar_expect_tx(&full[s], g_bytes + x_bytes);
            tma_load_2d(st, &tm_g, &full[s], l * W + c * 32, b * p.c_out);
            tma_load_2d(st + g_bytes, &tm_x, &full[s], l * W + c * 32, b * p.c_in);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(p.m64 ? 64 : 128, N, false, false);
      const uint32_t idesc2 = idesc_tf32(p.m64 ? 64 : 128, 32, false, false);
      // The issuing thread is the critical resource of this kernel (8 MMAs per 16 KB stage): descriptors are built once,
      // per MMA only the 14-bit start-address field moves (>> 4 units: +2 per 32-byte K step)
      const uint64_t d0 = smem_desc_sw128(smem_u32(stage0), 0, 1024);
      const uint32_t hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
      const uint32_t stage16 = stage_bytes >> 4, g16 = g_bytes >> 4;
      const uint32_t tab_lo0 = (uint32_t)smem_desc_sw128(smem_u32(smem), 0, 1024);
      (void)idesc2;
      (void)tab_lo0;
      // weight-gradient MMAs only: the first-DFT-stage MMAs of the same stages are issued by warp 3 (the single issuing
      // thread was the critical resource: 8 MMAs + barrier polling per 16 KB stage)
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int64_t gi = first; gi < groups; gi += stride) {
       for (int rr = 0; rr < kWgGroup; ++rr) {
        if (gi * kWgGroup + rr >= p.rows) break;
        for (int c = 0; c < p.chunks_per_row; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_lo = lo0 + (uint32_t)s * stage16;  // A: 128 rows from the g box on
          const uint32_t b_lo = a_lo + g16;                   // B: x box (+ ones rows)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_tf32_lohi(tmem_base, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc, acc);
            acc = 1;
          }
          mma_commit(&empty[s]);
          if (++s == S) { s = 0; ph ^= 1; }
        }
       }
      }
      mma_commit(done);
    }
  } else if (warp == 3) {
    // second MMA issuer (fused only): X1[(b, o, l), 0:32] += g rows . table chunk, per image row its own accumulator slice
    if (lane == 0 && fused) {
      const uint32_t idesc2 = idesc_tf32(p.m64 ? 64 : 128, 32, false, false);
      const uint64_t d0 = smem_desc_sw128(smem_u32(stage0), 0, 1024);
      const uint32_t hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
      const uint32_t stage16 = stage_bytes >> 4;
      const uint32_t tab_lo0 = (uint32_t)smem_desc_sw128(smem_u32(smem), 0, 1024);
      mbar_wait(tabfull, 0);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int64_t gi = first; gi < groups; gi += stride, ++it) {
        const int a2 = it & 1;
        mbar_wait(&x1empty[a2], (((uint32_t)it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int rr = 0; rr < kWgGroup; ++rr) {
          if (gi * kWgGroup + rr >= p.rows) break;
          // the rows of a group land in consecutive 32-column slices of one accumulator tile: ONE drain per group
          const uint32_t d2 = tmem_base + 128u + (uint32_t)(a2 * kWgGroup * 32 + rr * 32);
          for (int c = 0; c < p.chunks_per_row; ++c) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a_lo = lo0 + (uint32_t)s * stage16;
            // same A rows (g rows 0..c_out-1 are the ones kept), B = table chunk [32 k x 32 w] of this pixel chunk
            const uint32_t t_lo = tab_lo0 + (uint32_t)c * (4096 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32_lohi(d2, a_lo + k * 2, hi, t_lo + k * 2, hi, idesc2, (c | k) ? 1u : 0u);
            mma_commit(&empty[s]);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        }
        mma_commit(&x1full[a2]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    if (fused && (q < 2 || p.m64)) {
      // drain the r2c accumulator of every row: TMEM lane = A row = output channel o (M = 128: lanes 0..63 hold the g rows;
      // M = 64: 16 rows in the first 16 lanes of every lane quarter)
      int it = 0;
      const int o = p.m64 ? (lane < 16 ? q * 16 + lane : 1 << 20) : q * 32 + lane;
      for (int64_t gi = first; gi < groups; gi += stride, ++it) {
        const int a2 = it & 1;
        mbar_wait(&x1full[a2], ((uint32_t)it >> 1) & 1);
        tc_fence_after();
        // all rows of the group are loaded before the single wait: one TMEM round trip per group instead of per row
        float v[kWgGroup][32];
#pragma unroll
        for (int rr = 0; rr < kWgGroup; ++rr)
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)(a2 * kWgGroup * 32 + rr * 32), v[rr]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x1empty[a2]);
        if (o < p.c_out) {
#pragma unroll
          for (int rr = 0; rr < kWgGroup; ++rr) {
            const int64_t r = gi * kWgGroup + rr;
            if (r < p.rows) {
              const int64_t row = ((r / p.lead) * p.c_out + o) * p.lead + r % p.lead;
              float4* dst = (float4*)(p.x1 + row * 32);
              if (p.round_x1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[rr][j] = round_tf32(v[rr][j]);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[rr][4 * j], v[rr][4 * j + 1], v[rr][4 * j + 2], v[rr][4 * j + 3]);
            }
          }
        }
      }
    }
    if (first < groups) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int o = p.m64 ? (lane < 16 ? q * 16 + lane : 1 << 20) : q * 32 + lane;
      const bool vec_ok = (p.c_in & 3) == 0 && ((uintptr_t)p.gw & 15) == 0;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (o < p.c_out) {
          if (c0 + 16 <= p.c_in && vec_ok) {
            // 128-bit reductions: 148 CTAs add into the same c_out x c_in words, a quarter of the L2 atomic operations
            float4* dst = (float4*)(p.gw + (int64_t)o * p.c_in + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dst + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int i = c0 + j;
              if (i < p.c_in) atomicAdd(p.gw + (int64_t)o * p.c_in + i, v[j]);
              else if (i == p.c_in_pad && p.gb) atomicAdd(p.gb + o, v[j]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// small prep kernel: rounded (and optionally transposed) copy of the bypass weight
__global__ void prep_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int co, int ci, int transpose) {
  pdl_enter();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= co * ci) return;
  int o = idx / ci, i = idx % ci;
  float v = round_tf32(w[idx]);
  if (transpose) out[i * co + o] = v;
  else out[idx] = v;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// plan-side state and launchers
// ------------------------------------------------------------------------------------------
struct TcPlan {
  bool r2c_ok = false, slab_ok = false, wgrad_ok = false;
  int sms = 148;
  float* tab_r2c[2] = {nullptr, nullptr};  // [32][s_D]  (n-major rows, K contiguous): forward / adjoint-of-inverse
  float* tab_c2r[2] = {nullptr, nullptr};  // [s_D][32]  (w rows, K contiguous): inverse / adjoint-of-forward
  CUtensorMap tm_tab_r2c[2], tm_tab_c2r[2];
  int slab_stages = 0;
};

static TcPlan* tcp(const pdeqx_spectral_plan* p) { return (TcPlan*)p->tc; }

int tc_plan_init(pdeqx_spectral_plan* p) {
  const pdeqx_spectral_desc& d = p->desc;
  p->tc = nullptr;
  if (d.precision != PDEQX_PREC_TF32) return PDEQX_OK;
  int dev = 0, major = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return PDEQX_OK;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10 || !tc::encode_tiled_fn()) return PDEQX_OK;
  const int D = d.ndim, sD = d.spatial[D - 1], mD = p->mlast;
  if (2 * mD != 32) return PDEQX_OK;  // 16 (padded) modes on the contiguous axis = 32 accumulator columns
  if (sD % 32 != 0) return PDEQX_OK;
  TcPlan* t = new TcPlan();
  t->sms = sms;
  // tables in tensor-core layout, rounded to tf32 (round-to-nearest: no truncation bias from the constants)
  std::vector<float> h((size_t)32 * sD);
  std::vector<float> src((size_t)32 * sD);
  const float* dev_src_r2c[2] = {p->t_r2c_fwd, p->t_r2c_bwd};  // [s_D][32]
  const float* dev_src_c2r[2] = {p->t_c2r_inv, p->t_c2r_bwd};  // [32][s_D]
  int rc = PDEQX_OK;
  for (int v = 0; v < 2 && rc == PDEQX_OK; ++v) {
    if (cudaMemcpy(src.data(), dev_src_r2c[v], src.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = PDEQX_ERR_CUDA;
    for (int n = 0; n < 32; ++n)
      for (int k = 0; k < sD; ++k) h[(size_t)n * sD + k] = tc::host_round_tf32(src[(size_t)k * 32 + n]);
    if (cudaMalloc((void**)&t->tab_r2c[v], h.size() * 4) != cudaSuccess ||
        cudaMemcpy(t->tab_r2c[v], h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = PDEQX_ERR_CUDA;
    if (cudaMemcpy(src.data(), dev_src_c2r[v], src.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = PDEQX_ERR_CUDA;
    // W >= 128: [s_D rows][32 k]. W < 128 (128 % W == 0): the slab unit spans rpu = 128 / W image rows; tile c (of rpu)
    // is [128 pixels][32 k] with the table in the rows of image row c and zeros elsewhere (block-diagonal operand).
    const int rpu = sD >= 128 ? 1 : (128 % sD == 0 ? 128 / sD : 0);
    std::vector<float> hc((size_t)(sD >= 128 ? sD : rpu * 128) * 32, 0.f);
    if (sD >= 128) {
      for (int w = 0; w < sD; ++w)
        for (int k = 0; k < 32; ++k) hc[(size_t)w * 32 + k] = tc::host_round_tf32(src[(size_t)k * sD + w]);
    } else {
      for (int c = 0; c < rpu; ++c)
        for (int w = 0; w < sD; ++w)
          for (int k = 0; k < 32; ++k)
            hc[((size_t)c * 128 + (size_t)c * sD + w) * 32 + k] = tc::host_round_tf32(src[(size_t)k * sD + w]);
    }
    if (hc.empty()) hc.resize(32, 0.f);
    if (cudaMalloc((void**)&t->tab_c2r[v], hc.size() * 4) != cudaSuccess ||
        cudaMemcpy(t->tab_c2r[v], hc.data(), hc.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = PDEQX_ERR_CUDA;
    if (rc == PDEQX_OK) {
      uint64_t dims[2] = {(uint64_t)sD, 32};
      uint64_t str[1] = {(uint64_t)sD * 4};
      uint32_t box[2] = {32, 32};
      rc = tc::make_tensor_map(&t->tm_tab_r2c[v], t->tab_r2c[v], 2, dims, str, box);
    }
    if (rc == PDEQX_OK && (sD % 128 == 0 || rpu > 0)) {
      uint64_t dims[2] = {32, (uint64_t)(sD >= 128 ? sD : rpu * 128)};
      uint64_t str[1] = {32 * 4};
      uint32_t box[2] = {32, 128};
      rc = tc::make_tensor_map(&t->tm_tab_c2r[v], t->tab_c2r[v], 2, dims, str, box);
    }
  }
  if (rc != PDEQX_OK) {
    set_error("tcgen05 plan setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    p->tc = t;
    tc_plan_free(p);
    return rc;
  }
  t->r2c_ok = (sD <= 1024);
  // slab: both directions (c_in <-> c_out) must fit
  auto ch_ok = [](int ci, int co) { return ci % 8 == 0 && ci >= 8 && ci <= 128 && co % 32 == 0 && co >= 32 && co <= 128; };
  const bool w_ok = sD % 128 == 0 || (sD >= 32 && sD < 128 && 128 % sD == 0 && p->lead % (128 / sD) == 0);
  if (w_ok && ch_ok(d.c_in, d.c_out) && ch_ok(d.c_out, d.c_in)) {
    // every variant the block uses must fit two load stages: forward / data-gradient (2 tiles) and the
    // pre-activation-cotangent pass (3 tiles)
    t->slab_ok = tc::slab_smem(sD, d.c_in, d.c_out, true, 2, 2).total <= 227 * 1024 &&
                 tc::slab_smem(sD, d.c_out, d.c_in, true, 2, 2).total <= 227 * 1024 &&
                 tc::slab_smem(sD, d.c_in, d.c_out, true, 2, 3).total <= 227 * 1024;
    t->slab_stages = 2;
  }
  t->wgrad_ok = d.c_in % 16 == 0 && d.c_out % 8 == 0 && d.c_in <= 112 && d.c_out <= 128;
  p->tc = t;
  return PDEQX_OK;
}

void tc_plan_free(pdeqx_spectral_plan* p) {
  TcPlan* t = tcp(p);
  if (!t) return;
  for (int v = 0; v < 2; ++v) {
    cudaFree(t->tab_r2c[v]);
    cudaFree(t->tab_c2r[v]);
  }
  delete t;
  p->tc = nullptr;
}

int tc_stage_mask(const pdeqx_spectral_plan* p) {
  TcPlan* t = tcp(p);
  if (!t || !g_fast_paths.load(std::memory_order_relaxed)) return 0;
  return (t->r2c_ok ? 1 : 0) | (t->slab_ok ? 2 : 0) | (t->wgrad_ok ? 4 : 0) | (tc_mode_gemm_available(p) ? 8 : 0);
}

bool tc_r2c_available(const pdeqx_spectral_plan* p) {
  return tcp(p) && tcp(p)->r2c_ok && g_fast_paths.load(std::memory_order_relaxed);
}

int tc_r2c(const pdeqx_spectral_plan* p, const float* x, bool adjoint_table, int64_t rows, float* out, cudaStream_t s) {
  TcPlan* t = tcp(p);
  const int D = p->desc.ndim, sD = p->desc.spatial[D - 1];
  PDEQX_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "tc_r2c: pointers must be 16-byte aligned");
  CUtensorMap tm_x;
  uint64_t dims[2] = {(uint64_t)sD, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)sD * 4};
  uint32_t box[2] = {32, 128};
  PDEQX_TRY(tc::make_tensor_map(&tm_x, x, 2, dims, str, box));
  tc::R2cParams rp{};
  rp.rows = rows;
  rp.tiles = ceil_div(rows, 128);
  rp.kchunks = sD / 32;
  rp.n_valid = 32;
  rp.out_ld = 32;
  rp.out = out;
  rp.round_out = D >= 2 ? 1 : 0;  // the next stage is a tcgen05 c2c stage (it would truncate); D == 1 feeds the fp32 channel mix
  const uint32_t tab_bytes = (uint32_t)rp.kchunks * 4096;
  int stages = (int)((227 * 1024 - 2048 - tab_bytes) / 16384);
  if (stages > 10) stages = 10;
  PDEQX_CHECK_ARG(stages >= 2, "tc_r2c: table does not leave room for the pipeline");
  rp.stages = stages;
  const size_t smem = tab_bytes + (size_t)stages * 16384 + 1024 + 1024;
  static DeviceOnce attr;      
  if (attr.first_use()) {
    PDEQX_CUDA(cudaFuncSetAttribute(tc::r2c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                
  }
  const int grid = (int)(rp.tiles < t->sms ? rp.tiles : t->sms);
  PDEQX_CUDA(launch_pdl(tc::r2c_kernel, dim3(grid), dim3(tc::kThreadsR2c), smem, s, tm_x, t->tm_tab_r2c[adjoint_table ? 1 : 0], rp));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

bool tc_slab_available(const pdeqx_spectral_plan* p, bool has_bypass) {
  (void)has_bypass;
  return tcp(p) && tcp(p)->slab_ok && g_fast_paths.load(std::memory_order_relaxed);
}

// y[B, c_o, lead, W] = epi( c2r(z)[table_sel] + Wmat . x + bias ), Wmat = bypass_w (c_o x c_i) or its transpose
// (bypass_transposed: bypass_w is [c_i_of_this_call... see spectral.cu) ; mode 1: y = aux * act'(.)
int tc_slab_ex(const pdeqx_spectral_plan* p, float* w_prep, const float* x, const float* z, int table_sel, const float* bypass_w,
               bool bypass_transposed, const float* bypass_b, int act, int mode, const float* aux, int c_in, int c_out,
               float* y, cudaStream_t s, const float* aux_w, float* red_w, float* red_b, const float* byp_u, const float* byp_a) {
  TcPlan* t = tcp(p);
  const pdeqx_spectral_desc& d = p->desc;
  const int D = d.ndim, sD = d.spatial[D - 1];
  const int64_t lead = p->lead;
  PDEQX_CHECK_ARG(mode != 2 || (aux && aux_w), "tc_slab: mode 2 needs the cotangent factors");
  PDEQX_CHECK_ARG(mode != 3 || (aux && red_w && act == PDEQX_ACT_IDENTITY), "tc_slab: mode 3 needs the field, an output and no activation");
  PDEQX_CHECK_ARG(mode != 4 || (y && aux_w), "tc_slab: mode 4 needs the projection weight and the output field");
  PDEQX_CHECK_ARG(mode != 5 || (aux && aux_w && red_w && c_out <= 64), "tc_slab: mode 5 needs the cotangent factors, an output and c_out <= 64");
  float* y_field = y;  // modes 3 / 4 store no tile: the tile tensor map is a placeholder
  if (mode == 3) y = red_w;
  PDEQX_CHECK_ARG(((uintptr_t)z & 15) == 0 && ((uintptr_t)y & 15) == 0 && (!x || ((uintptr_t)x & 15) == 0),
                  "tc_slab: pointers must be 16-byte aligned");
  const bool bypass = bypass_w != nullptr;
  PDEQX_CHECK_ARG(!bypass || w_prep, "tc_slab: the bypass needs c_in * c_out floats of scratch");
  CUtensorMap tm_x, tm_z, tm_w;
  {
    uint64_t dims[3] = {32, (uint64_t)lead, (uint64_t)d.batch * c_out};
    uint64_t str[2] = {32 * 4, (uint64_t)lead * 32 * 4};
    uint32_t box[3] = {32, 1, (uint32_t)c_out};
    PDEQX_TRY(tc::make_tensor_map(&tm_z, z, 3, dims, str, box));
  }
  if (bypass) {
    uint64_t dims[3] = {(uint64_t)sD, (uint64_t)lead, (uint64_t)d.batch * c_in};
    uint64_t str[2] = {(uint64_t)sD * 4, (uint64_t)lead * sD * 4};
    uint32_t box[3] = {32, 1, (uint32_t)c_in};
    PDEQX_TRY(tc::make_tensor_map(&tm_x, x, 3, dims, str, box, tc::TM_SW128_ATOM32));
    // rounded copy of the weight as [c_out rows][c_in] (K contiguous)
    const int n = c_in * c_out;
    // bypass_w is stored [rows_w][cols_w]; not transposed: [c_out][c_in]; transposed: stored [c_in][c_out]
    PDEQX_CUDA(launch_pdl(tc::prep_weight_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, bypass_w, w_prep, bypass_transposed ? c_in : c_out,
                                                          bypass_transposed ? c_out : c_in, bypass_transposed ? 1 : 0));
    PDEQX_LAUNCHED();
    uint64_t wd[2] = {(uint64_t)c_in, (uint64_t)c_out};
    uint64_t ws[1] = {(uint64_t)c_in * 4};
    uint32_t wb[2] = {32, (uint32_t)c_out};
    PDEQX_TRY(tc::make_tensor_map(&tm_w, w_prep, 2, wd, ws, wb));
  } else {
    tm_x = tm_z;
    tm_w = tm_z;
  }
  CUtensorMap tm_y, tm_aux;
  {
    uint64_t dims[3] = {(uint64_t)sD, (uint64_t)lead, (uint64_t)d.batch * c_out};
    uint64_t str[2] = {(uint64_t)sD * 4, (uint64_t)lead * sD * 4};
    uint32_t box[3] = {32, 1, (uint32_t)(c_out / 4)};  // one epilogue warp's piece
    if (mode == 4) tm_y = tm_z;  // (never used)
    else PDEQX_TRY(tc::make_tensor_map(&tm_y, y, 3, dims, str, box));
    if (mode == 1) {
      PDEQX_CHECK_ARG(aux && ((uintptr_t)aux & 15) == 0, "tc_slab: aux must be a 16-byte aligned device pointer");
      PDEQX_TRY(tc::make_tensor_map(&tm_aux, aux, 3, dims, str, box));
    } else {
      tm_aux = tm_y;
    }
  }
  tc::SlabParams sp{};
  sp.lead = (int)lead;
  sp.W = sD;
  sp.parts = 4;  // c_out is a multiple of 32: 8, 16 or 32 channels per epilogue warp
  sp.rpu = sD >= 128 ? 1 : 128 / sD;
  sp.wt_per_row = sD >= 128 ? sD / 128 : 1;
  sp.c_in = c_in;
  sp.c_out = c_out;
  sp.has_bypass = bypass ? 1 : 0;
  sp.mode = mode;
  sp.act = act;
  sp.round_out = 0;
  sp.units = (int64_t)d.batch * (lead / sp.rpu) * sp.wt_per_row;
  sp.bias = bypass_b;
  sp.aux = aux;
  sp.aux_w = aux_w;
  sp.byp_u = byp_u;
  sp.byp_a = byp_a;
  sp.red_w = red_w;
  sp.red_b = red_b;
  sp.y = mode == 4 ? y_field : y;
  // mode 1 keeps three epilogue tiles (cotangent prefetch distance 2), modes 3 / 4 store no tile and keep none; the load
  // pipeline gets what is left
  const int epi_bufs = mode == 1 ? 3 : (mode == 3 || mode == 4) ? 0 : 2;
  int stages = 0;
  for (int st = 6; st >= 2; --st)
    if (tc::slab_smem(sD, c_in, c_out, bypass, st, epi_bufs).total <= 227 * 1024) { stages = st; break; }
  PDEQX_CHECK_ARG(stages >= 2, "tc_slab: shared memory does not fit two load stages (c_in %d, c_out %d)", c_in, c_out);
  sp.stages = stages;
  tc::SlabSmem L = tc::slab_smem(sD, c_in, c_out, bypass, sp.stages, epi_bufs);
  int grid = (int)(sp.units < t->sms ? sp.units : t->sms);
  grid = grid / sp.wt_per_row * sp.wt_per_row;  // every CTA keeps one w-tile of the table (units stride by the grid size)
  PDEQX_CHECK_ARG(grid >= sp.wt_per_row, "tc_slab: grid smaller than the number of w-tiles");
  PDEQX_CHECK_ARG(mode == 0 || mode == 3 || mode == 4 || act != PDEQX_ACT_IDENTITY, "tc_slab: modes 1, 2 and 5 need an activation");
  const CUtensorMap& tab = t->tm_tab_c2r[table_sel];
  const bool byp1 = byp_u != nullptr;
  PDEQX_CHECK_ARG(!byp1 || (byp_a && !bypass && mode != 3), "tc_slab: a rank-one bypass replaces the bypass weight");
#define PDEQX_SLAB_LAUNCH_B(M, A, B1)                                                                                     \
  do {                                                                                                                    \
    static DeviceOnce attr;                                                                                                   \
    if (attr.first_use()) {                                                                                                          \
      PDEQX_CUDA(cudaFuncSetAttribute(tc::slab_kernel<M, A, B1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                                                                                                                          \
    }                                                                                                                     \
    PDEQX_CUDA(launch_pdl(tc::slab_kernel<M, A, B1>, dim3(grid), dim3(tc::kThreadsSlab), L.total, s, tm_x, tm_z, tab, tm_w, tm_y, tm_aux, sp, L)); \
  } while (0)
#define PDEQX_SLAB_LAUNCH(M, A)                 \
  do {                                          \
    if (byp1) PDEQX_SLAB_LAUNCH_B(M, A, true);  \
    else PDEQX_SLAB_LAUNCH_B(M, A, false);      \
  } while (0)
  if (mode == 3) PDEQX_SLAB_LAUNCH_B(3, PDEQX_ACT_IDENTITY, false);
  else if (mode == 0 && act == PDEQX_ACT_IDENTITY) PDEQX_SLAB_LAUNCH(0, PDEQX_ACT_IDENTITY);
  else if (mode == 0 && act == PDEQX_ACT_RELU) PDEQX_SLAB_LAUNCH(0, PDEQX_ACT_RELU);
  else if (mode == 0) PDEQX_SLAB_LAUNCH(0, PDEQX_ACT_GELU_TANH);
  else if (mode == 1 && act == PDEQX_ACT_RELU) PDEQX_SLAB_LAUNCH(1, PDEQX_ACT_RELU);
  else if (mode == 1) PDEQX_SLAB_LAUNCH(1, PDEQX_ACT_GELU_TANH);
  else if (mode == 2 && act == PDEQX_ACT_RELU) PDEQX_SLAB_LAUNCH(2, PDEQX_ACT_RELU);
  else if (mode == 2) PDEQX_SLAB_LAUNCH(2, PDEQX_ACT_GELU_TANH);
  else if (mode == 4 && act == PDEQX_ACT_IDENTITY) PDEQX_SLAB_LAUNCH(4, PDEQX_ACT_IDENTITY);
  else if (mode == 4 && act == PDEQX_ACT_RELU) PDEQX_SLAB_LAUNCH(4, PDEQX_ACT_RELU);
  else if (mode == 4) PDEQX_SLAB_LAUNCH(4, PDEQX_ACT_GELU_TANH);
  else if (act == PDEQX_ACT_RELU) PDEQX_SLAB_LAUNCH(5, PDEQX_ACT_RELU);
  else PDEQX_SLAB_LAUNCH(5, PDEQX_ACT_GELU_TANH);
#undef PDEQX_SLAB_LAUNCH_B
#undef PDEQX_SLAB_LAUNCH
  PDEQX_LAUNCHED();
  {
    // algorithmic bytes of this launch: every tensor it reads or writes, once
    const double n_pix = (double)lead * sD, B = (double)d.batch;
    double bytes = 4.0 * B * c_out * (double)lead * 32;                       // z
    if (bypass) bytes += 4.0 * B * c_in * n_pix + 4.0 * c_in * c_out;         // x slab (+ bypass weight)
    if (mode == 1) bytes += 4.0 * B * c_out * n_pix;                          // cotangent tile
    if (mode == 2 || mode == 3 || mode == 5) bytes += 4.0 * B * n_pix;        // single-channel cotangent factor / field
    if (byp1) bytes += 4.0 * B * n_pix;                                       // single-channel field behind the lifting
    if (mode == 4) bytes += 4.0 * B * n_pix;                                  // projected field
    else if (mode != 3) bytes += 4.0 * B * c_out * n_pix;                     // output tile
    bytes += 4.0 * 32 * sD + 4.0 * c_out;                                     // table + bias
    timed_bytes(bytes);
  }
  return PDEQX_OK;
}

int tc_slab(const pdeqx_spectral_plan* p, float* w_prep, const float* x, const float* z, int table_sel, const float* bypass_w,
            bool bypass_transposed, const float* bypass_b, int act, float* y, cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  const int c_in = bypass_transposed ? d.c_out : d.c_in, c_out = bypass_transposed ? d.c_in : d.c_out;
  return tc_slab_ex(p, w_prep, x, z, table_sel, bypass_w, bypass_transposed, bypass_b, act, 0, nullptr, c_in, c_out, y, s, nullptr, nullptr, nullptr, nullptr, nullptr);
}


// ---- chain: data gradient of block `p` fused with the pre-activation cotangent of the block `up` in front of it ----
static int chain_mode() {  // PDEQX_CHAIN=0 switches the fusion off (A/B measurements)
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("PDEQX_CHAIN");
    env = e ? atoi(e) : 1;
  }
  return env;
}

bool tc_chain_available(const pdeqx_spectral_plan* p, const pdeqx_spectral_plan* up) {
  if (!p || !up || !chain_mode() || !tc_slab_available(p, true) || !tc_slab_available(up, true)) return false;
  const pdeqx_spectral_desc &a = p->desc, &b = up->desc;
  if (a.ndim != b.ndim || a.batch != b.batch || p->lead != up->lead) return false;
  for (int i = 0; i < a.ndim; ++i)
    if (a.spatial[i] != b.spatial[i]) return false;
  const int C = a.c_in;
  if (a.c_out != C || b.c_in != C || b.c_out != C || (C != 32 && C != 64)) return false;
  return tc::chain_smem(a.spatial[a.ndim - 1], C, 2).total <= 227 * 1024;
}

// out[B, C, lead, W] = ( c2r_adj(z1) + w_k^T . g ) * act'( c2r_inv(z_up) + w_up . x_up + bias_up )
//   byp_u / byp_a (x_up == null): the upstream block's bypass term is rank one, byp_a[c] * byp_u[b, pixel] (+ bias_up)
int tc_chain(const pdeqx_spectral_plan* p, float* w_prep, float* w_prep2, const float* g, const float* z1, const float* w_k, const pdeqx_spectral_plan* up,
             const float* x_up, const float* z_up, const float* w_up, const float* bias_up, int act, float* out, cudaStream_t s,
             const float* byp_u, const float* byp_a) {
  TcPlan* t = tcp(p);
  TcPlan* tu = tcp(up);
  const pdeqx_spectral_desc& d = p->desc;
  const int D = d.ndim, sD = d.spatial[D - 1], C = d.c_in;
  const int64_t lead = p->lead;
  const bool byp1 = x_up == nullptr;
  PDEQX_CHECK_ARG(tc_chain_available(p, up), "tc_chain: the two blocks do not qualify for the fused pass");
  PDEQX_CHECK_ARG(g && z1 && w_k && z_up && out && (byp1 ? (byp_u && byp_a) : (w_up != nullptr)), "tc_chain: null argument");
  PDEQX_CHECK_ARG(act == PDEQX_ACT_RELU || act == PDEQX_ACT_GELU_TANH, "tc_chain: the upstream block needs an activation");
  PDEQX_CHECK_ARG((((uintptr_t)g | (uintptr_t)z1 | (uintptr_t)z_up | (uintptr_t)out | (uintptr_t)x_up) & 15) == 0,
                  "tc_chain: pointers must be 16-byte aligned");
  CUtensorMap tm_g, tm_z1, tm_x2, tm_z2, tm_w1, tm_w2, tm_y;
  {
    uint64_t dims[3] = {32, (uint64_t)lead, (uint64_t)d.batch * C};
    uint64_t str[2] = {32 * 4, (uint64_t)lead * 32 * 4};
    uint32_t box[3] = {32, 1, (uint32_t)C};
    PDEQX_TRY(tc::make_tensor_map(&tm_z1, z1, 3, dims, str, box));
    PDEQX_TRY(tc::make_tensor_map(&tm_z2, z_up, 3, dims, str, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)sD, (uint64_t)lead, (uint64_t)d.batch * C};
    uint64_t str[2] = {(uint64_t)sD * 4, (uint64_t)lead * sD * 4};
    uint32_t box[3] = {32, 1, (uint32_t)C};
    PDEQX_TRY(tc::make_tensor_map(&tm_g, g, 3, dims, str, box, tc::TM_SW128_ATOM32));
    if (!byp1) PDEQX_TRY(tc::make_tensor_map(&tm_x2, x_up, 3, dims, str, box, tc::TM_SW128_ATOM32));
    else tm_x2 = tm_g;
    uint32_t ybox[3] = {32, 1, (uint32_t)(C / 4)};
    PDEQX_TRY(tc::make_tensor_map(&tm_y, out, 3, dims, str, ybox));
  }
  {
    const int n = C * C;
    PDEQX_CUDA(launch_pdl(tc::prep_weight_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, w_k, w_prep, C, C, 1));  // w_k^T: [c_in rows][c_out]
    PDEQX_LAUNCHED();
    uint64_t wd[2] = {(uint64_t)C, (uint64_t)C};
    uint64_t ws[1] = {(uint64_t)C * 4};
    uint32_t wb[2] = {32, (uint32_t)C};
    PDEQX_TRY(tc::make_tensor_map(&tm_w1, w_prep, 2, wd, ws, wb));
    if (!byp1) {
      PDEQX_CUDA(launch_pdl(tc::prep_weight_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, w_up, w_prep2, C, C, 0));
      PDEQX_LAUNCHED();
      PDEQX_TRY(tc::make_tensor_map(&tm_w2, w_prep2, 2, wd, ws, wb));
    } else {
      tm_w2 = tm_w1;
    }
  }
  tc::ChainParams cp{};
  cp.lead = (int)lead;
  cp.W = sD;
  cp.C = C;
  cp.rpu = sD >= 128 ? 1 : 128 / sD;
  cp.wt_per_row = sD >= 128 ? sD / 128 : 1;
  cp.units = (int64_t)d.batch * (lead / cp.rpu) * cp.wt_per_row;
  cp.bias = bias_up;
  cp.byp_u = byp_u;
  cp.byp_a = byp_a;
  int stages = 0;
  for (int st = 4; st >= 2; --st)
    if (tc::chain_smem(sD, C, st).total <= 227 * 1024) { stages = st; break; }
  PDEQX_CHECK_ARG(stages >= 2, "tc_chain: shared memory does not fit two load stages");
  cp.stages = stages;
  tc::ChainSmem L = tc::chain_smem(sD, C, stages);
  int grid = (int)(cp.units < t->sms ? cp.units : t->sms);
  grid = grid / cp.wt_per_row * cp.wt_per_row;
  PDEQX_CHECK_ARG(grid >= cp.wt_per_row, "tc_chain: grid smaller than the number of w-tiles");
#define PDEQX_CHAIN_LAUNCH(A, B1)                                                                                          \
  do {                                                                                                                     \
    static DeviceOnce attr;                                                                                                    \
    if (attr.first_use()) {                                                                                                           \
      PDEQX_CUDA(cudaFuncSetAttribute(tc::chain_kernel<A, B1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));  \
                                                                                                                           \
    }                                                                                                                      \
    PDEQX_CUDA(launch_pdl(tc::chain_kernel<A, B1>, dim3(grid), dim3(tc::kThreadsSlab), L.total, s, tm_g, tm_z1, tm_x2, tm_z2, t->tm_tab_c2r[1], \
                          tu->tm_tab_c2r[0], tm_w1, tm_w2, tm_y, cp, L));                                                    \
  } while (0)
  if (act == PDEQX_ACT_GELU_TANH) {
    if (byp1) PDEQX_CHAIN_LAUNCH(PDEQX_ACT_GELU_TANH, true);
    else PDEQX_CHAIN_LAUNCH(PDEQX_ACT_GELU_TANH, false);
  } else {
    if (byp1) PDEQX_CHAIN_LAUNCH(PDEQX_ACT_RELU, true);
    else PDEQX_CHAIN_LAUNCH(PDEQX_ACT_RELU, false);
  }
#undef PDEQX_CHAIN_LAUNCH
  PDEQX_LAUNCHED();
  {
    const double n_pix = (double)lead * sD, B = (double)d.batch;
    double bytes = 2 * 4.0 * B * C * (double)lead * 32;           // both z tensors
    bytes += 4.0 * B * C * n_pix;                                 // g_k
    bytes += byp1 ? 4.0 * B * n_pix : 4.0 * B * C * n_pix;        // upstream input (or the single-channel field behind it)
    bytes += 4.0 * B * C * n_pix;                                 // g_{k-1}
    bytes += 2 * 4.0 * 32 * sD + 2 * 4.0 * C * C + 4.0 * C;
    timed_bytes(bytes);
  }
  return PDEQX_OK;
}

bool tc_wgrad_available(const pdeqx_spectral_plan* p, int64_t n_pix) {
  return tcp(p) && tcp(p)->wgrad_ok && n_pix % 32 == 0 && g_fast_paths.load(std::memory_order_relaxed);
}
bool tc_wgrad_r2c_available(const pdeqx_spectral_plan* p) {
  const pdeqx_spectral_desc& d = p->desc;
  return tcp(p) && tcp(p)->wgrad_ok && tcp(p)->r2c_ok && d.c_out <= 64 && d.spatial[d.ndim - 1] <= 1024 &&
         g_fast_paths.load(std::memory_order_relaxed);
}

// gw [c_out, c_in] and gb [c_out] must be zeroed by the caller. x1_out (optional): the first stage of the
// adjoint-of-inverse DFT of g, [B * c_out * lead][32] -- the same pass over g that feeds the weight gradient
int tc_wgrad(const pdeqx_spectral_plan* p, const float* g, const float* x, int64_t n_pix, float* gw, float* gb, cudaStream_t s,
             float* x1_out, bool rank1) {
  TcPlan* t = tcp(p);
  const pdeqx_spectral_desc& d = p->desc;
  const int sD = d.spatial[d.ndim - 1];
  PDEQX_CHECK_ARG(n_pix == p->lead * sD, "tc_wgrad: n_pix does not match the plan");
  PDEQX_CHECK_ARG(!x1_out || (d.c_out <= 64 && p->mlast == 16 && t->r2c_ok), "tc_wgrad: fused r2c needs c_out <= 64, 16 modes");
  CUtensorMap tm_g, tm_x;
  {
    uint64_t dims[2] = {(uint64_t)n_pix, (uint64_t)d.batch * d.c_out};
    uint64_t str[1] = {(uint64_t)n_pix * 4};
    uint32_t box[2] = {32, (uint32_t)d.c_out};
    PDEQX_TRY(tc::make_tensor_map(&tm_g, g, 2, dims, str, box));
  }
  // rank1: x is the single-channel field [B, 1, pixels] in front of a 1 -> C lifting; gw is [c_out] (one column)
  const int c_in = rank1 ? 1 : d.c_in, c_in_pad = rank1 ? 16 : d.c_in;
  {
    uint64_t dims[2] = {(uint64_t)n_pix, (uint64_t)d.batch * c_in};
    uint64_t str[1] = {(uint64_t)n_pix * 4};
    uint32_t box[2] = {32, (uint32_t)c_in_pad};
    PDEQX_TRY(tc::make_tensor_map(&tm_x, x, 2, dims, str, box));
  }
  tc::WgradParams wp{};
  wp.c_out = d.c_out;
  wp.c_in = c_in;
  wp.c_in_pad = c_in_pad;
  wp.chunks_per_row = sD / 32;
  wp.lead = (int)p->lead;
  wp.rows = (int64_t)d.batch * p->lead;
  wp.gw = gw;
  wp.gb = gb;
  wp.x1 = x1_out;
  wp.round_x1 = d.ndim >= 2 ? 1 : 0;
  static const bool m64_on = [] { const char* e = getenv("PDEQX_WG_M64"); return !(e && e[0] == '0'); }();  // (0: M = 128, A = [g ; x])
  wp.m64 = (m64_on && d.c_out <= 64) ? 1 : 0;
  const uint32_t tab_bytes = x1_out ? (uint32_t)wp.chunks_per_row * 4096 : 0;
  const uint32_t stage_bytes = (uint32_t)(d.c_out + c_in_pad) * 128 + 2048;
  int stages = (int)((227 * 1024 - 16384 - 2048 - (int)tab_bytes) / (int)stage_bytes);
  if (stages > 12) stages = 12;
  PDEQX_CHECK_ARG(stages >= 2, "tc_wgrad: shared memory does not fit two stages");
  wp.stages = stages;
  const size_t smem = tab_bytes + (size_t)stages * stage_bytes + 16384 + 1024 + 1024;
  static DeviceOnce attr;      
  if (attr.first_use()) {
    PDEQX_CUDA(cudaFuncSetAttribute(tc::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                
  }
  const int64_t wgroups = ceil_div(wp.rows, tc::kWgGroup);
  const int grid = (int)(wgroups < t->sms ? wgroups : t->sms);
  PDEQX_CUDA(launch_pdl(tc::wgrad_kernel, dim3(grid), dim3(tc::kThreadsR2c), smem, s, tm_g, tm_x, t->tm_tab_r2c[1], wp));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx
