// PhysicsConv on tcgen05 (TF32): boundary-aware 2-D k x k cross-correlation as an implicit GEMM over image rows.
//
//   replaces jnp.pad + lax.conv_general_dilated + bias (pdequinox/conv/_conv.py:169-219) for the hot-path case:
//   2-D, stride 1, groups 1, odd kernel k (SAME padding p = d (k-1)/2), any dilation d, W == 128, C_i, C_o multiples
//   of 32 up to 64 -- the ClassicResNet / DilatedResNet layers of BASELINE.json configs[2].
//
// One unit = one output row h of one sample for one half of the output channels:
//     D[w', (tw, o)] = sum_{th, i} x[b, i, rowmap(h + th d - p), w'] * W[o, i, th, tw]      (M = 128, N = k*C_o/2, K = k*C_i)
//     y[b, o, h, w]  = act( bias[o] + residual + sum_tw D[colmap(w + tw d - p), (tw, o)] )
//   * the ROW taps are different TMA boxes (any boundary mode is a row-index map; dirichlet rows are skipped);
//     consecutive units walk rows with stride d, so two of the three input rows of a unit are still resident in a
//     3-slot shared-memory ring: ONE new row (C_i x 512 B) is fetched per unit;
//   * the COLUMN taps stay separate accumulator columns and are shifted in the epilogue, where periodic / dirichlet /
//     neumann is again just an index map (wrap / drop / mirror) -- no padded copy, no halo exchange;
//   * the x row is the MN-major A operand (pixels contiguous): TMA SWIZZLE_128B_ATOM_32B + UMMA SWIZZLE_128B_BASE32B;
//     the re-laid-out, tf32-rounded weights of the CTA's channel half stay resident in shared memory (K-major B).
// The data gradient of periodic / dirichlet convolutions is the same kernel with flipped, transposed weights.
// At C = 64, k = 3 the layer sits at the TF32 ridge (144 flop/B): tensor time ~ HBM time.
#include "conv_plan.h"
#include "tc_common.cuh"

namespace pdeqx {
namespace tc {

constexpr int kConvThreads = 640;  // warps 0-2: TMA / MMA / TMEM alloc; warps 4-19: epilogue (4 per TMEM lane quarter)
constexpr int kConvOcnt = 8;       // output channels per epilogue warp (NOH = 32 channels per CTA / 4 warps per quarter)
constexpr int kConvTmemCols = 512;
constexpr int kAcc = 4;  // accumulator stages in TMEM (4 x 96 columns): the MMA issuer runs up to 4 units ahead of the epilogue
constexpr int kMaxK = 3;
constexpr int kRing = 4;    // input rows resident per CTA: 3 in use + 1 in flight
constexpr int kMaxDil = 8;  // column taps reach at most 8 columns into the neighbouring warp's columns

struct ConvParams {
  int H, W, c_in, c_out, k, d, p, boundary;
  int batch, seg, segs_per_image;
  int act, has_residual, has_bias;
  int64_t items;  // batch * 2 halves * segs_per_image
  const float* bias;
  float* y;          // pair kernel: output / residual as plain pointers (its epilogue stores from registers)
  const float* res;
  const float* mask;  // optional [B, c_out, H, W]: the result is zeroed where mask <= 0 (relu' of the layer in front, applied
                      // in the data-gradient pass that produces its cotangent instead of in a pass of its own)
  // GroupNorm (groups = 1) folded into the pair kernel (template parameter GN, see conv_rows_pair_kernel):
  const float* gn_stats;  // GN 1: [B][2] (mean, rstd) of the conv input -- the accumulator is scaled by rstd_b in the epilogue
  const float* bias_b;    // GN 1: [B][c_out] per-sample epilogue constant (bias + what the norm's shift contributes)
  float* parts;           // GN 1: [B][slots][2] partial sums (y, y^2) of the output = the NEXT norm's statistics (may be NULL);
                          // GN 2: [B][slots][2] partial sums (gamma v, gamma v h) of the data gradient v (the norm's reverse pass)
  int slots;              // partial sums per sample: items per sample x 16 epilogue warps
};

struct ConvSmem {
  uint32_t w_off, w_th_bytes, w_chunk_bytes, ring_off, slot_bytes, box_bytes, tile_off, tile_bytes, xbuf_off, bar_off, total;
};

static ConvSmem conv_smem(int c_in, int c_out, int k) {
  ConvSmem s{};
  const uint32_t N = (uint32_t)k * (c_out / 2);
  uint32_t off = 0;
  s.w_off = off;
  s.w_chunk_bytes = N * 128;
  s.w_th_bytes = (uint32_t)((c_in + 31) / 32) * s.w_chunk_bytes;
  off += (uint32_t)k * s.w_th_bytes;
  off = (off + 1023) & ~1023u;
  s.ring_off = off;
  s.box_bytes = (uint32_t)c_in * 128;
  s.slot_bytes = 4 * s.box_bytes;
  off += kRing * s.slot_bytes;
  s.tile_off = off;
  s.tile_bytes = (uint32_t)(c_out / 2) * 512;
  off += s.tile_bytes;  // one output / residual tile (the ring needs the rest of the 227 KB)
  s.xbuf_off = off;     // edge-column exchange: [2 taps][16 warps][kMaxDil columns][8 channels] floats
  off += 2 * 16 * 8 * kMaxDil * 4 + 64;  // (+ one all-zero column slot: the source of dropped dirichlet taps, pair kernel)
  s.bar_off = off;
  off += 1024;
  s.total = off + 1024;
  return s;
}

__device__ __forceinline__ int conv_map(int q, int n, int mode) {
  if (q >= 0 && q < n) return q;
  if (mode == PDEQX_PAD_WRAP) {
    int r = q % n;
    return r < 0 ? r + n : r;
  }
  if (mode == PDEQX_PAD_REFLECT) {
    const int per = 2 * (n - 1);
    int r = q % per;
    if (r < 0) r += per;
    return r < n ? r : per - r;
  }
  if (mode == PDEQX_PAD_EDGE) return q < 0 ? 0 : n - 1;
  return -1;
}

// Work decomposition. Rows are walked in stride-d chains (r0, r0 + d, r0 + 2d, ...): the three input rows of
// output row h = r0 + j d are the "virtual" chain rows v = j-1, j, j+1 (actual row = rowmap(r0 + v d)), so consecutive
// units share two of them. Virtual row v lives in ring slot (v + 1) % 3: a unit loads only v = j+1 (all three at the
// start of a segment) and hands v = j-1 back after its first tap. The TMA producer, the MMA issuer and the epilogue
// warps all derive this schedule from the item index alone -- a few integer ops per unit, nothing is communicated.
//   item = ((b * d + r0) * segs_per_chain + seg) * 2 + half
struct ConvItem {
  int b, half, r0, j0, count;
};
__device__ __forceinline__ ConvItem conv_item(const ConvParams& p, int64_t item) {
  ConvItem it;
  it.half = (int)(item & 1);
  int64_t r = item >> 1;
  const int seg = (int)(r % p.segs_per_image);  // segments per chain
  r /= p.segs_per_image;
  it.r0 = (int)(r % p.d);
  it.b = (int)(r / p.d);
  it.j0 = seg * p.seg;
  const int len = (p.H - it.r0 + p.d - 1) / p.d;
  int c = len - it.j0;
  it.count = c < 0 ? 0 : (c > p.seg ? p.seg : c);
  return it;
}

template <int ACT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_res, ConvParams p, ConvSmem L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + L.bar_off);
  uint64_t* full = bars;         // [kRing] ring slots
  uint64_t* empty = bars + 4;    // [kRing]
  uint64_t* tfull = bars + 8;    // [kAcc] accumulators
  uint64_t* tempty = bars + 12;  // [kAcc]
  uint64_t* wfull = bars + 16;   // weights
  uint64_t* rfull = bars + 17;   // residual tile
  uint32_t* tmem_ptr = (uint32_t*)(bars + 19);
  float* bias_s = (float*)(bars + 20);  // [c_out/2]
  uint64_t* rfull_w = bars + 40;        // [16 epilogue warps] residual piece landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NOH = p.c_out / 2, N = 3 * NOH;
  const int half_of_cta = (int)(blockIdx.x & 1);  // items alternate halves and the grid is even: a CTA keeps its half

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_res);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);
    }
    mbar_init(rfull, 1);
    mbar_init(wfull, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&rfull_w[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kConvTmemCols>(tmem_ptr);
  if (warp == 3)
    for (int i = lane; i < NOH; i += 32) bias_s[i] = p.has_bias ? p.bias[half_of_cta * NOH + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int kchunks = (p.c_in + 31) / 32;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(wfull, 3 * L.w_th_bytes);
      for (int t = 0; t < 3; ++t)
        for (int c = 0; c < kchunks; ++c)
          tma_load_2d(smem + L.w_off + t * L.w_th_bytes + c * L.w_chunk_bytes, &tm_w, wfull, c * 32,
                      (half_of_cta * 3 + t) * N);
      uint32_t pe = 0, used = 0;  // per-slot parity of the empty barrier / "has been filled before"
      for (int64_t item = first; item < p.items; item += stride) {
        const ConvItem it = conv_item(p, item);
        if (it.count == 0) continue;
        // chain rows j0-1 .. j0+count of this segment, in order; runs ahead as far as the ring allows
        for (int v = it.j0 - 1; v <= it.j0 + it.count; ++v) {
          const int row = conv_map(it.r0 + v * p.d, p.H, p.boundary);
          if (row < 0) continue;  // dirichlet: the tap is skipped
          const int s = (v + 1) & (kRing - 1);
          if (used & (1u << s)) {
            mbar_wait_backoff(&empty[s], (pe >> s) & 1u);
            pe ^= 1u << s;
          }
          used |= 1u << s;
          uint8_t* dst = smem + L.ring_off + (size_t)s * L.slot_bytes;
          mbar_expect_tx(&full[s], L.slot_bytes);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * L.box_bytes, &tm_x, &full[s], c * 32, row, it.b * p.c_in);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ONE thread issues every tcgen05.mma of the CTA, so its instruction stream is the critical path once the
    // memory side keeps up: descriptors are (constant high word, low word = base + offset), the k loop is unrolled,
    // and row validity needs no division.
    // The whole warp runs the loop convergently (so the descriptor arithmetic stays warp-uniform) and one elected
    // lane issues the tcgen05 instructions.
    {
      const uint32_t idesc = idesc_tf32(128, N, /*a_mn_major=*/true, /*b_mn_major=*/false);
      mbar_wait(wfull, 0);
      int a = 0;
      uint32_t aph = 0, pf = 0;
      const uint64_t adesc0 = smem_desc_sw128_atom32(smem_u32(smem + L.ring_off), L.box_bytes, 512);
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem + L.w_off), 0, 1024);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t slot16 = L.slot_bytes >> 4, wth16 = L.w_th_bytes >> 4, chunk16 = L.w_chunk_bytes >> 4;
      const bool zeros = p.boundary == PDEQX_PAD_ZEROS;
      for (int64_t item = first; item < p.items; item += stride) {
        const ConvItem it = conv_item(p, item);
        for (int jj = 0; jj < it.count; ++jj) {
          const int j = it.j0 + jj;
          const bool last = jj == it.count - 1;
          const int h = it.r0 + j * p.d;
          mbar_wait(&tempty[a], aph ^ 1);
          tc_fence_after();
          const uint32_t dcol = tmem_base + (uint32_t)(a * N);
          uint32_t acc = 0;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            if (zeros && ((t == 0 && h - p.d < 0) || (t == 2 && h + p.d >= p.H))) continue;  // dirichlet row tap
            const int s = (j + t) & (kRing - 1);  // slot of chain row v = j - 1 + t
            if (jj == 0 || t == 2) {              // first use of this row
              mbar_wait(&full[s], (pf >> s) & 1u);
              pf ^= 1u << s;
              tc_fence_after();
            }
            const uint32_t a_lo = a_lo0 + (uint32_t)s * slot16;
            const uint32_t b_lo = b_lo0 + (uint32_t)t * wth16;
            if (elect_one()) {
            if (p.c_in == 64) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                mma_tf32_lohi(dcol, a_lo + k * 64, a_hi, b_lo + (k >> 2) * chunk16 + (k & 3) * 2, b_hi, idesc, acc);
                acc = 1;
              }
            } else {
              const int ks = p.c_in / 8;
              for (int k = 0; k < ks; ++k) {
                mma_tf32_lohi(dcol, a_lo + k * 64, a_hi, b_lo + (k >> 2) * chunk16 + (k & 3) * 2, b_hi, idesc, acc);
                acc = 1;
              }
            }
            if (t == 0 || last) mma_commit(&empty[s]);  // v = j-1 is never needed again; at a segment end nothing is
            }
            acc = 1;
            __syncwarp();
          }
          if (elect_one()) mma_commit(&tfull[a]);
          __syncwarp();
          if (++a == kAcc) { a = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // This stage is latency bound (TMEM load -> neighbour exchange -> shared store), so it gets 16 warps: warp (og, q)
    // owns 8 output channels x the 32 columns of TMEM lane quarter q. Its output piece is its own swizzled 1 KB TMA box
    // (stored, and for a residual pre-loaded, by the warp itself); the only cross-warp step is the exchange of the d edge
    // columns between the four lane-quarter warps of one channel group (one 128-thread named barrier).
    const int q = warp & 3;
    const int og = (warp - 4) >> 2;            // which 8-channel group of the half
    constexpr int ocnt = kConvOcnt;            // channels per warp
    const int wid = og * 4 + q;                // epilogue warp index 0..15
    const int wq = q * 32 + lane;              // this thread's source column w'
    int a = 0;
    uint32_t aph = 0, eph = 0, par = 0;
    // per-thread constants of the column gather: output column wq reads D_0 at colmap(wq - d) and D_2 at colmap(wq + d)
    const int c0 = conv_map(wq - p.d, p.W, p.boundary), c2 = conv_map(wq + p.d, p.W, p.boundary);
    const bool use0 = c0 >= 0, use2 = c2 >= 0;
    const bool in0 = use0 && (c0 >> 5) == q, in2 = use2 && (c2 >> 5) == q;
    const int l0 = c0 & 31, l2 = c2 & 31;
    const int d = p.d;
    // the exchange slots are double-buffered by unit parity when d <= 4 (two halves of the kMaxDil entries): one barrier
    // per unit; wider dilations keep a second barrier that protects the buffer against the next unit's writes
    const bool dbuf = d <= kMaxDil / 2;
    float bias_r[ocnt];
#pragma unroll
    for (int jx = 0; jx < ocnt; ++jx) bias_r[jx] = bias_s[og * ocnt + jx];
    uint8_t* piece_p = smem + L.tile_off + (size_t)wid * 1024;  // [8 channels][32 w] fp32, 128-byte rows, SWIZZLE_128B
    const uint32_t piece = smem_u32(piece_p);
    const uint32_t lane_b = (uint32_t)lane * 4;
    float* xb = (float*)(smem + L.xbuf_off);
    const uint32_t xbs = smem_u32(xb);
    // where this lane's out-of-warp sources sit in the exchange buffer (own slot 0 if it has none: the value is not used)
    const uint32_t far0_base = xbs + (((0 * 16 + (uint32_t)(og * 4 + (use0 && !in0 ? (c0 >> 5) : q))) * kMaxDil +
                                       (uint32_t)(use0 && !in0 ? (c0 & 31) - (32 - d) : 0)) * ocnt) * 4;
    const uint32_t far2_base = xbs + (((1 * 16 + (uint32_t)(og * 4 + (use2 && !in2 ? (c2 >> 5) : q))) * kMaxDil +
                                       (uint32_t)(use2 && !in2 ? (c2 & 31) : 0)) * ocnt) * 4;
    for (int64_t item = first; item < p.items; item += stride) {
      const ConvItem it = conv_item(p, item);
      const int ch0 = it.b * p.c_out + it.half * NOH + og * ocnt;
      if (p.has_residual && lane == 0 && it.count > 0) {  // first residual piece of the item
        tma_store_wait_read<0>();
        mbar_expect_tx(&rfull_w[wid], 1024);
        tma_load_3d(piece_p, &tm_res, &rfull_w[wid], q * 32, it.r0 + it.j0 * p.d, ch0);
      }
      for (int jj = 0; jj < it.count; ++jj) {
        const int h = it.r0 + (it.j0 + jj) * p.d;
        mbar_wait(&tfull[a], aph);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * N + og * ocnt);
        // the three column taps of this thread's SOURCE column w' = wq (8 channels each), then the accumulator is free
        float v0[ocnt], v1[ocnt], v2[ocnt];
        tmem_ld8(t0, v0);
        tmem_ld8(t0 + (uint32_t)NOH, v1);
        tmem_ld8(t0 + (uint32_t)(2 * NOH), v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);
        // Output column w gathers D_0[colmap(w - d)] + D_1[w] + D_2[colmap(w + d)]. Sources inside the warp's own 32 columns
        // come by shuffle; the d columns next to a warp edge (and the wrapped ones) go through a small exchange buffer.
        // Exchange layout [tap][warp][column slot][8 channels]: a column's 8 channels are 32 contiguous bytes, so an edge lane
        // writes / a neighbour reads them with two 128-bit accesses (the shared-memory pipe is shared with the tensor core's
        // operand fetch, which alone takes half of it: 16 scalar loads + 16 scalar stores per unit and warp were as much again)
        const uint32_t xo = dbuf ? par * (kMaxDil / 2) : 0u;
        if (lane >= 32 - d) {  // my last d columns of D_0
          const uint32_t a0 = xbs + (((0 * 16 + (uint32_t)wid) * kMaxDil + xo + (uint32_t)(lane - (32 - d))) * ocnt) * 4;
          st_shared_v4(a0, v0);
          st_shared_v4(a0 + 16, v0 + 4);
        }
        if (lane < d) {  // my first d columns of D_2
          const uint32_t a2 = xbs + (((1 * 16 + (uint32_t)wid) * kMaxDil + xo + (uint32_t)lane) * ocnt) * 4;
          st_shared_v4(a2, v2);
          st_shared_v4(a2 + 16, v2 + 4);
        }
        if (p.has_residual) {
          mbar_wait(&rfull_w[wid], eph);
          eph ^= 1;
        } else {
          if (lane == 0) tma_store_wait_read<0>();  // this warp's previous store has drained its piece
          __syncwarp();
        }
        named_bar_sync(1 + og, 128);
        // branch-free gather: every lane loads a (valid) exchange-buffer address and selects
        const uint32_t far0 = far0_base + xo * (ocnt * 4), far2 = far2_base + xo * (ocnt * 4);
        float f0[ocnt], f2[ocnt], res[ocnt];
        ld_shared_v4(far0, f0);
        ld_shared_v4(far0 + 16, f0 + 4);
        ld_shared_v4(far2, f2);
        ld_shared_v4(far2 + 16, f2 + 4);
        if (p.has_residual) {
#pragma unroll
          for (int jx = 0; jx < ocnt; ++jx) res[jx] = ld_shared_f32(piece + (uint32_t)jx * 128 + (lane_b ^ ((uint32_t)jx << 4)));
        }
        if (!dbuf) named_bar_sync(1 + og, 128);  // everyone has read the exchange slots before the next unit overwrites them
#pragma unroll
        for (int jx = 0; jx < ocnt; ++jx) {
          const float s0 = __shfl_sync(0xffffffffu, v0[jx], l0);
          const float s2 = __shfl_sync(0xffffffffu, v2[jx], l2);
          float r = v1[jx] + bias_r[jx];
          r += use0 ? (in0 ? s0 : f0[jx]) : 0.f;
          r += use2 ? (in2 ? s2 : f2[jx]) : 0.f;
          if (p.has_residual) r += res[jx];
          if (ACT == PDEQX_ACT_RELU) r = fmaxf(r, 0.f);
          else if (ACT == PDEQX_ACT_GELU_TANH) r = gelu_tanh(r);
          if (p.mask && __ldg(p.mask + (((int64_t)ch0 + jx) * p.H + h) * p.W + wq) <= 0.f) r = 0.f;
          // the next consumer is a tcgen05 stage (conv / data gradient / filter gradient), which would TRUNCATE fp32 to
          // tf32 (a coherent bias of about -2^-12 per layer): round to nearest here
          r = round_tf32(r);
          st_shared_f32(piece + (uint32_t)jx * 128 + (lane_b ^ ((uint32_t)jx << 4)), r);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tm_y, (const void*)piece_p, q * 32, h, ch0);
          tma_store_commit();
          if (p.has_residual && jj + 1 < it.count) {  // next residual piece once the store has read this one
            tma_store_wait_read<0>();
            mbar_expect_tx(&rfull_w[wid], 1024);
            tma_load_3d(piece_p, &tm_res, &rfull_w[wid], q * 32, h + p.d, ch0);
          }
        }
        par ^= 1;
        if (++a == kAcc) { a = 0; aph ^= 1; }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kConvTmemCols>(tmem_base);
  }
}

// ==========================================================================================
// CTA-pair variant (cta_group::2): the same implicit GEMM with ALL output channels per unit.
// ==========================================================================================
// The one-CTA kernel above keeps one half of the filter (73 KB) next to the 4-row input ring, so a unit is N = 96 wide:
// per loaded row it has only 24 x 56 cycles of tensor work, with ONE 32 KB row in flight per SM -- it is bound by the
// latency of that single outstanding row (Little: 148 x 32 KB / ~1.2 us = 3.9 TB/s of operand traffic, which it needs
// twice because both channel halves read every x row). Here two CTAs of one TPC form a pair: each keeps ITS half of the
// filter and ITS OWN image rows (M = 256 = 128 pixels of each CTA's row), and one tcgen05.mma.cta_group::2 of N = 192
// reads the filter halves from both shared memories. Per loaded row a CTA now has 24 x 96 cycles of tensor work, every x
// row is fetched once, and the same 4-slot ring covers the load latency twice over.
//   * both CTAs load their own rows / filter half with cta_group::2 TMA loads that count on the LEADER's barriers;
//   * the leader's elected lane issues every MMA and commits (multicast) to the ring / accumulator barriers of both CTAs;
//   * the epilogue warps of both CTAs drain their own TMEM (16 warps x 2 passes of 8 channels), store straight from
//     registers (a warp-wide store per channel is one 128-byte line) and release the accumulator on the leader's barrier;
//   * dirichlet row taps are zero rows (out-of-range TMA coordinate) instead of skipped MMAs: the two CTAs work on
//     different rows and must issue the same instruction stream.
constexpr int kAcc2 = 2;  // accumulator stages: 2 x 192 columns

// GN = 0: plain. GN = 1: forward pass with a GroupNorm (groups = 1) in front folded in -- the filter carries gamma (prepared
// weights), the epilogue computes act(rstd_b * D + bias_b[b, o]) and sums y, y^2 per sample for the norm BEHIND this layer.
// GN = 2: data gradient whose result v feeds that norm's reverse pass -- bias_s holds gamma, p.mask the norm's input h (read,
// not applied): the epilogue sums gamma v and gamma v h per sample (pdequinox/blocks/_dilated_res_block.py:141-151).
template <int ACT, int GN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_rows_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_res, ConvParams p, ConvSmem L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + L.bar_off);
  uint64_t* full = bars;         // [kRing] ring slots (the leader's count both CTAs' bytes)
  uint64_t* empty = bars + 4;    // [kRing]
  uint64_t* tfull = bars + 8;    // [kAcc2] accumulators
  uint64_t* tempty = bars + 12;  // [kAcc2] (the leader's: 16 epilogue warps of each CTA)
  uint64_t* wfull = bars + 16;   // weights (the leader's)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 19);
  float* bias_s = (float*)(bars + 20);  // [c_out = 64]
  uint64_t* xbar = bars + 56;           // [4 channel groups][2] edge columns of a pass are in the exchange buffer

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NOH = p.c_out / 2, NH = 3 * NOH, N = 2 * NH;  // 32, 96 (per CTA), 192 (per MMA)
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_res);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAcc2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 32);
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&xbar[i], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<kConvTmemCols>(tmem_ptr);
  if (warp == 3) {
    for (int i = lane; i < p.c_out; i += 32) bias_s[i] = p.has_bias ? p.bias[i] : (GN == 2 ? 1.f : 0.f);  // (GN 2: gamma)
    if (lane < 16) ((float*)(smem + L.xbuf_off + 2 * 16 * 8 * kMaxDil * 4))[lane] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // pair q of the grid takes item pairs q, q + pairs, ...; the CTA of rank r works on item 2 * pair_index + r
  const int64_t first = (int64_t)(blockIdx.x >> 1) * 2 + rank, stride = (int64_t)(gridDim.x >> 1) * 2;
  const int kchunks = (p.c_in + 31) / 32;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      if (rank == 0) mbar_expect_tx(wfull, 2 * 3 * L.w_th_bytes);
      for (int t = 0; t < 3; ++t)
        for (int c = 0; c < kchunks; ++c)
          tma_load_2d_pair(smem + L.w_off + t * L.w_th_bytes + c * L.w_chunk_bytes, &tm_w, wfull, c * 32, ((int)rank * 3 + t) * NH);
      uint32_t pe = 0, used = 0;
      for (int64_t item = first; item < p.items; item += stride) {
        const ConvItem it = conv_item(p, item * 2);
        // the count + 2 chain rows of the item, in order; slots are numbered from the item's first row so that both CTAs of
        // the pair (different segments) walk the ring identically
        for (int i = 0; i <= it.count + 1; ++i) {
          const int row = conv_map(it.r0 + (it.j0 - 1 + i) * p.d, p.H, p.boundary);  // -1 (dirichlet halo): out of range = zeros
          const int s = i & (kRing - 1);
          if (used & (1u << s)) {
            mbar_wait_backoff(&empty[s], (pe >> s) & 1u);
            pe ^= 1u << s;
          }
          used |= 1u << s;
          uint8_t* dst = smem + L.ring_off + (size_t)s * L.slot_bytes;
          if (rank == 0) mbar_expect_tx(&full[s], 2 * L.slot_bytes);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_3d_pair(dst + c * L.box_bytes, &tm_x, &full[s], c * 32, row, it.b * p.c_in);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t idesc = idesc_tf32(256, N, /*a_mn_major=*/true, /*b_mn_major=*/false);
      mbar_wait(wfull, 0);
      int a = 0;
      uint32_t aph = 0, pf = 0;
      const uint64_t adesc0 = smem_desc_sw128_atom32(smem_u32(smem + L.ring_off), L.box_bytes, 512);
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem + L.w_off), 0, 1024);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t slot16 = L.slot_bytes >> 4, wth16 = L.w_th_bytes >> 4, chunk16 = L.w_chunk_bytes >> 4;
      const int ks = p.c_in / 8;
      for (int64_t item = first; item < p.items; item += stride) {
        const ConvItem it = conv_item(p, item * 2);
        for (int jj = 0; jj < it.count; ++jj) {
          const bool last = jj == it.count - 1;
          mbar_wait(&tempty[a], aph ^ 1);
          tc_fence_after();
          const uint32_t dcol = tmem_base + (uint32_t)(a * N);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const int s = (jj + t) & (kRing - 1);  // slot of the item's row jj + t (= chain row j - 1 + t)
            if (jj == 0 || t == 2) {               // first use of this row
              mbar_wait(&full[s], (pf >> s) & 1u);
              pf ^= 1u << s;
              tc_fence_after();
            }
            const uint32_t a_lo = a_lo0 + (uint32_t)s * slot16;
            const uint32_t b_lo = b_lo0 + (uint32_t)t * wth16;
            if (elect_one()) {
              for (int k = 0; k < ks; ++k)
                mma_tf32_lohi_pair(dcol, a_lo + k * 64, a_hi, b_lo + (k >> 2) * chunk16 + (k & 3) * 2, b_hi, idesc, (t | k) ? 1u : 0u);
              if (t == 0 || last) mma_commit_pair(&empty[s]);  // v = j-1 is never needed again; at a segment end nothing is
            }
            __syncwarp();
          }
          if (elect_one()) mma_commit_pair(&tfull[a]);
          __syncwarp();
          if (++a == kAcc2) { a = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own rows) =====================
    // warp (og, q) owns 16 output channels x the 32 columns of TMEM lane quarter q, drained in two passes of 8 channels;
    // a pass is the one-CTA kernel's unit epilogue (neighbour exchange, 1 KB swizzled piece, own TMA store).
    const int q = warp & 3;
    const int og = (warp - 4) >> 2;
    constexpr int ocnt = kConvOcnt;
    const int wid = og * 4 + q;
    const int wq = q * 32 + lane;
    int a = 0;
    uint32_t aph = 0, par = 0;
    const int c0 = conv_map(wq - p.d, p.W, p.boundary), c2 = conv_map(wq + p.d, p.W, p.boundary);
    const bool use0 = c0 >= 0, use2 = c2 >= 0;
    const bool in0 = use0 && (c0 >> 5) == q, in2 = use2 && (c2 >> 5) == q;
    const int l0 = c0 & 31, l2 = c2 & 31;
    const int d = p.d;
    const bool dbuf = d <= kMaxDil / 2;
    const int64_t plane = (int64_t)p.H * p.W;
    const uint32_t xbs = smem_u32(smem + L.xbuf_off);
    // where this lane's out-of-warp sources sit in the exchange buffer: a neighbour's slot, the all-zero slot for a dropped
    // (dirichlet) tap, or -- value unused -- the own slot 0 if the source is inside the warp (it comes by shuffle)
    const uint32_t zero_slot = xbs + 2 * 16 * 8 * kMaxDil * 4;
    const uint32_t far0_base = !use0 ? zero_slot
                                     : xbs + (((0 * 16 + (uint32_t)(og * 4 + (!in0 ? (c0 >> 5) : q))) * kMaxDil +
                                               (uint32_t)(!in0 ? (c0 & 31) - (32 - d) : 0)) * ocnt) * 4;
    const uint32_t far2_base = !use2 ? zero_slot
                                     : xbs + (((1 * 16 + (uint32_t)(og * 4 + (!in2 ? (c2 >> 5) : q))) * kMaxDil +
                                               (uint32_t)(!in2 ? (c2 & 31) : 0)) * ocnt) * 4;
    const uint32_t xo_b = dbuf ? (kMaxDil / 2) * (ocnt * 4) : 0u;  // byte offset of the second exchange buffer half
    const uint32_t xo0_b = use0 ? xo_b : 0u, xo2_b = use2 ? xo_b : 0u;  // (the zero slot has no second half)
    const uint32_t bias_sa = smem_u32(bias_s);
    const uint32_t bias_w = smem_u32(smem + L.tile_off) + (uint32_t)wid * 64;  // GN 1: this warp's per-sample constants
    uint32_t xph = 0;
    for (int64_t item = first; item < p.items; item += stride) {
      const ConvItem it = conv_item(p, item * 2);
      const int chb = it.b * p.c_out + og * 2 * ocnt;  // first of this warp's 16 channels
      float rstd = 1.f, sum_a = 0.f, sum_b = 0.f;
      if (GN == 1) {
        // the sample's epilogue constants of this warp's 16 channels: fetched once per item into the warp's own 64 bytes of
        // the (otherwise unused) output-tile region, read per pass like the shared bias of the plain kernel
        rstd = __ldg(p.gn_stats + 2 * it.b + 1);
        __syncwarp();
        if (lane < 2 * ocnt) st_shared_f32(bias_w + (uint32_t)lane * 4, __ldg(p.bias_b + (size_t)it.b * p.c_out + og * 2 * ocnt + lane));
        __syncwarp();
      }
      for (int jj = 0; jj < it.count; ++jj) {
        const int h = it.r0 + (it.j0 + jj) * p.d;
        // GN 2: the norm's input at this lane's 16 output elements is requested BEFORE the wait for the accumulator, so the
        // global-load latency overlaps the unit's MMAs instead of sitting on the epilogue's chain of two passes
        float pre_m[2][ocnt];
        if (GN == 2) {
#pragma unroll
          for (int ps = 0; ps < 2; ++ps)
#pragma unroll
            for (int jx = 0; jx < ocnt; ++jx)
              pre_m[ps][jx] = __ldg(p.mask + (((int64_t)chb + ps * ocnt + jx) * p.H + h) * p.W + wq);
        }
        mbar_wait(&tfull[a], aph);
        tc_fence_after();
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int cg = og * 2 + ps;  // 8-channel group 0..7: filter half cg >> 2, group cg & 3 inside the half
          const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * N + (cg >> 2) * NH + (cg & 3) * ocnt);
          // this lane's 8 output elements: pixel (h, wq) of 8 consecutive channel planes; a warp-wide store per channel is one
          // full 128-byte line, so the pass needs no staging tile, no async-proxy fence and no wait on an earlier TMA store
          const int64_t off0 = (((int64_t)chb + ps * ocnt) * p.H + h) * p.W + wq;
          float res[ocnt], msk[ocnt];
          if (p.has_residual) {
#pragma unroll
            for (int jx = 0; jx < ocnt; ++jx) res[jx] = __ldg(p.res + off0 + (int64_t)jx * plane);
          }
          if (GN == 2) {
#pragma unroll
            for (int jx = 0; jx < ocnt; ++jx) msk[jx] = pre_m[ps][jx];
          } else if (p.mask) {
#pragma unroll
            for (int jx = 0; jx < ocnt; ++jx) msk[jx] = __ldg(p.mask + off0 + (int64_t)jx * plane);
          }
          float v0[ocnt], v1[ocnt], v2[ocnt];
          tmem_ld8(t0, v0);
          tmem_ld8(t0 + (uint32_t)NOH, v1);
          tmem_ld8(t0 + (uint32_t)(2 * NOH), v2);
          tmem_ld_wait();
          if (ps == 1) {  // last read of this accumulator: hand it back to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[a]);
          }
#if defined(PDEQX_CONV_DIAG) && PDEQX_CONV_DIAG < 3  // timing experiments only: 1 = no neighbour exchange, 2 = no stores either
          {
#if PDEQX_CONV_DIAG == 1
#pragma unroll
            for (int jx = 0; jx < ocnt; ++jx) p.y[off0 + (int64_t)jx * plane] = v0[jx] + v1[jx] + v2[jx];
#else
            if (v0[0] + v1[1] + v2[2] == 123.456f) p.y[off0] = 0.f;
#endif
            par ^= 1;
            continue;
          }
#endif
          const uint32_t xo = dbuf ? par * (kMaxDil / 2) : 0u;
          if (lane >= 32 - d) {
            const uint32_t a0 = xbs + (((0 * 16 + (uint32_t)wid) * kMaxDil + xo + (uint32_t)(lane - (32 - d))) * ocnt) * 4;
            st_shared_v4(a0, v0);
            st_shared_v4(a0 + 16, v0 + 4);
          }
          if (lane < d) {
            const uint32_t a2 = xbs + (((1 * 16 + (uint32_t)wid) * kMaxDil + xo + (uint32_t)lane) * ocnt) * 4;
            st_shared_v4(a2, v2);
            st_shared_v4(a2 + 16, v2 + 4);
          }
          // the edge columns are published on an mbarrier (one arrive per warp) and the wait comes after the work that does
          // not need them -- in-warp shifts and the centre tap -- so the exchange latency hides behind it
          __syncwarp();
          if (lane == 0) mbar_arrive(&xbar[og * 2 + (int)par]);
          float bias_r[ocnt], acc[ocnt], s0[ocnt], s2[ocnt];
          if (GN == 1) {
            ld_shared_v4(bias_w + (uint32_t)ps * (ocnt * 4), bias_r);
            ld_shared_v4(bias_w + (uint32_t)ps * (ocnt * 4) + 16, bias_r + 4);
          } else {
            ld_shared_const_v4(bias_sa + (uint32_t)cg * (ocnt * 4), bias_r);
            ld_shared_const_v4(bias_sa + (uint32_t)cg * (ocnt * 4) + 16, bias_r + 4);
          }
#pragma unroll
          for (int jx = 0; jx < ocnt; ++jx) {
            s0[jx] = __shfl_sync(0xffffffffu, v0[jx], l0);
            s2[jx] = __shfl_sync(0xffffffffu, v2[jx], l2);
            if (GN == 1) acc[jx] = fmaf(v1[jx], rstd, bias_r[jx]);
            else if (GN == 2) acc[jx] = v1[jx];
            else acc[jx] = v1[jx] + bias_r[jx];
            if (p.has_residual) acc[jx] += res[jx];
          }
          mbar_wait(&xbar[og * 2 + (int)par], (xph >> par) & 1u);
          xph ^= 1u << par;
          const uint32_t far0 = far0_base + (par ? xo0_b : 0u), far2 = far2_base + (par ? xo2_b : 0u);
          // only the d lanes next to a warp edge (and dropped taps: the zero slot) read the exchange buffer: the shared-memory
          // pipe is shared with the tensor core's operand fetch, which is what bounds this kernel
          float f0[ocnt], f2[ocnt];
          if (!in0) {
            ld_shared_v4(far0, f0);
            ld_shared_v4(far0 + 16, f0 + 4);
          }
          if (!in2) {
            ld_shared_v4(far2, f2);
            ld_shared_v4(far2 + 16, f2 + 4);
          }
          __syncwarp();
          if (!dbuf) named_bar_sync(1 + og, 128);
#pragma unroll
          for (int jx = 0; jx < ocnt; ++jx) {
            float r;
            if (GN == 1) r = fmaf((in0 ? s0[jx] : f0[jx]) + (in2 ? s2[jx] : f2[jx]), rstd, acc[jx]);
            else r = acc[jx] + (in0 ? s0[jx] : f0[jx]) + (in2 ? s2[jx] : f2[jx]);
            if (ACT == PDEQX_ACT_RELU) r = fmaxf(r, 0.f);
            else if (ACT == PDEQX_ACT_GELU_TANH) r = gelu_tanh(r);
            if (GN == 2) {  // v stays fp32 (an elementwise pass reads it); msk = the norm's input h, bias_r = gamma
              const float t = bias_r[jx] * r;
              sum_a += t;
              sum_b = fmaf(t, msk[jx], sum_b);
              p.y[off0 + (int64_t)jx * plane] = r;
              continue;
            }
            if (p.mask && msk[jx] <= 0.f) r = 0.f;
#if defined(PDEQX_CONV_DIAG) && PDEQX_CONV_DIAG == 3  // timing experiment: full epilogue arithmetic, no stores
            if (r == 123.456f) p.y[off0] = r;
#else
            r = round_tf32(r);  // (read next by a tcgen05 stage: round, do not let it truncate)
            p.y[off0 + (int64_t)jx * plane] = r;
#endif
            if (GN == 1) {
              sum_a += r;
              sum_b = fmaf(r, r, sum_b);
            }
          }
          par ^= 1;
        }
        if (++a == kAcc2) { a = 0; aph ^= 1; }
      }
      if (GN != 0 && p.parts) {  // this warp's partial sums of the item (one sample): a slot of its own, reduced later
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          sum_a += __shfl_xor_sync(0xffffffffu, sum_a, off);
          sum_b += __shfl_xor_sync(0xffffffffu, sum_b, off);
        }
        if (lane == 0) {
          const int64_t local = item - (int64_t)it.b * (p.d * p.segs_per_image);
          float* dst = p.parts + (((int64_t)it.b * p.slots) + local * 16 + wid) * 2;
          dst[0] = sum_a;
          dst[1] = sum_b;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every MMA, remote arrive and TMEM read of the pair is done before either CTA lets go
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<kConvTmemCols>(tmem_base);
  }
}

// weights -> [half][th][n = tw * NOH + o_local][i], tf32-rounded. flip: data-gradient weights
// W'[i_as_out][o_as_in][th][tw] = W[o][i][k-1-th][k-1-tw]
// in_scale [c_in] (may be NULL): gamma of a GroupNorm in front of the convolution, folded into the filter (forward only)
__global__ void conv_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int c_out, int c_in, int k,
                                         int flip, const float* __restrict__ in_scale) {
  // (c_out, c_in) are the channel counts of THIS call (after the swap for the data gradient)
  const int NOH = c_out / 2;
  const int total = c_out * c_in * k * k;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int i = idx % c_in;
    int r = idx / c_in;
    const int ol = r % NOH;
    r /= NOH;
    const int tw = r % k;
    r /= k;
    const int th = r % k;
    const int half = r / k;
    const int o = half * NOH + ol;
    float v;
    if (!flip) v = w[((o * c_in + i) * k + th) * k + tw];
    else v = w[((i * c_out + o) * k + (k - 1 - th)) * k + (k - 1 - tw)];  // stored [c_in_call = orig out][c_out_call = orig in]
    if (in_scale) v *= __ldg(in_scale + i);
    out[idx] = round_tf32(v);
  }
}

// GroupNorm (groups = 1) in front of a periodic convolution, folded: conv(gamma_c (x - mu_b) rstd_b + beta_c) =
//   rstd_b conv(x; W gamma) + [ bias_o + sum_{c,t} W[o,c,t] beta_c - mu_b rstd_b sum_{c,t} W[o,c,t] gamma_c ]
// (every tap of a periodic -- or mirrored -- halo reads a real, normalised pixel, so the shift's contribution is a constant
// per sample and output channel). out[b][o] = the bracket. One block per o.
__global__ void __launch_bounds__(128)
conv_gn_bias_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ stats, float* __restrict__ out, int c_out,
                    int c_in, int taps, int batch) {
  __shared__ float red_a[4], red_b[4];
  const int o = blockIdx.x;
  float sb = 0.f, sg = 0.f;
  for (int i = threadIdx.x; i < c_in * taps; i += blockDim.x) {
    const int c = i / taps;
    const float wv = __ldg(w + (size_t)o * c_in * taps + i);
    sb = fmaf(wv, beta ? __ldg(beta + c) : 0.f, sb);
    sg += round_tf32(wv * (gamma ? __ldg(gamma + c) : 1.f));  // the filter as the tensor core reads it: the mean cancels exactly
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
    sg += __shfl_xor_sync(0xffffffffu, sg, off);
  }
  if ((threadIdx.x & 31) == 0) {
    red_a[threadIdx.x >> 5] = sb;
    red_b[threadIdx.x >> 5] = sg;
  }
  __syncthreads();
  sb = (red_a[0] + red_a[1]) + (red_a[2] + red_a[3]);
  sg = (red_b[0] + red_b[1]) + (red_b[2] + red_b[3]);
  const float b0 = (bias ? __ldg(bias + o) : 0.f) + sb;
  for (int b = threadIdx.x; b < batch; b += blockDim.x)
    out[(size_t)b * c_out + o] = b0 - __ldg(stats + 2 * b) * __ldg(stats + 2 * b + 1) * sg;
}

// ==========================================================================================
// Filter gradient on tcgen05:  gW[o,i,th,tw] = sum_{b,h,w} gy[b,o,h,w] * x[b,i,rowmap(h+(th-1)d), colmap(w+(tw-1)d)]
// ==========================================================================================
// Unit = one row (b, h); the contraction runs over the 128 columns of the row (K = w).
//   A (M = 128): column-SHIFTED copies of the gy row, stacked: group 1 = [shift for tw=0 (64 o) ; shift for tw=2 (64 o)],
//                group 2 = [centre (64 o) ; don't-care]. For d % 4 == 0 a shifted copy is just a TMA box whose start
//                column is moved by -s (out-of-range columns zero-filled: the zero-extended, "dirichlet in w" part; the d
//                wrapped / mirrored edge columns of periodic / neumann are added by conv_wgrad_edge_kernel). An unaligned
//                box column faults, so for the other dilations the two shifted operands are materialised once by
//                conv_shift_rows_kernel -- with the halo contributions already folded in, for every boundary mode.
//   B (N = 64):  the three x rows rowmap(h + (th-1) d) (K-major), from the same 3-row ring as the forward kernel.
//   D: 6 accumulators of 64 columns (2 groups x 3 row taps) that live in TMEM for the CTA's whole life; each CTA flushes
//      its partial sum to a scratch slab once, a second kernel reduces the slabs into gW (deterministic, no atomics).
struct WgParams {
  int H, W, c, d, boundary;
  int seg, segs_per_chain;
  int64_t items;  // batch * d * segs_per_chain
  float* scratch;  // [grid][3 tw][3 th][c i][c o]
  float* bias_part;  // [grid][c o] per-CTA sums of gy (TMEM-operand kernel; nullptr: no bias gradient wanted)
  // GroupNorm (groups = 1) in front of the convolution folded in: x is the norm's INPUT h, the gradient is taken against
  // gamma_c (h - mu_b) rstd_b + beta_c without that tensor existing -- the A operand is scaled by rstd_b on its way into
  // tensor memory, the rest is a rank-one correction in the reduce kernel (needs sum gy and sum mu_b rstd_b gy per o)
  const float* gn_stats;  // [B][2] mean, rstd (nullptr: plain filter gradient)
  float* m_part;          // [grid][c o] per-CTA sums of mu_b * (rstd_b gy as the tensor core reads it: tf32-rounded)
};

constexpr int kWgRing = 4;   // x rows resident (3 in use + 1 in flight)

struct WgItem {
  int b, r0, j0, count;
};
__device__ __forceinline__ WgItem wg_item(const WgParams& p, int64_t item) {
  WgItem it;
  int64_t r = item;
  const int seg = (int)(r % p.segs_per_chain);
  r /= p.segs_per_chain;
  it.r0 = (int)(r % p.d);
  it.b = (int)(r / p.d);
  it.j0 = seg * p.seg;
  const int len = (p.H - it.r0 + p.d - 1) / p.d;
  int c = len - it.j0;
  it.count = c < 0 ? 0 : (c > p.seg ? p.seg : c);
  return it;
}

// ------------------------------------------------------------------------------------------
// Filter gradient with the column-shifted operands built IN SHARED MEMORY (any dilation, any boundary mode).
// ------------------------------------------------------------------------------------------
// The kernel above needs the shifted copies of the gy row as TMA boxes, and a box column that is not 16-byte aligned
// faults -- so for d % 4 != 0 it reads two pre-shifted tensors that a separate pass had to write (3 extra passes over
// an activation-sized tensor). Here the gy row is loaded ONCE (raw, double buffered), and thirteen shifter warps
// write the two shifted copies of every 32-column chunk into a two-stage operand buffer -- with the wrapped / mirrored
// halo columns of periodic / neumann already folded in -- while the tensor core works on the previous chunk. The centre tap
// reads the raw row directly. HBM traffic: x and gy once.
struct WgsSmem {
  uint32_t raw_off, sh_off, ring_off, slot_bytes, box_bytes, bar_off, total;
};
constexpr int kWgsThreads = 512;
constexpr int kWgsShifters = kWgsThreads / 32 - 3;  // warps 3..15

static WgsSmem wgs_smem(int c) {
  WgsSmem s{};
  uint32_t off = 0;
  s.box_bytes = (uint32_t)c * 128;  // [c rows][32 w]
  s.raw_off = off;                  // 2 raw gy rows of 4 boxes; the box after a centre box is its (don't-care) upper half
  off += 2 * 4 * s.box_bytes;
  s.sh_off = off;                   // 2 stages of [shift tw=0 box][shift tw=2 box]
  off += 2 * 2 * s.box_bytes;
  s.ring_off = off;
  s.slot_bytes = 4 * s.box_bytes;
  off += kWgRing * s.slot_bytes;
  s.bar_off = off;
  off += 1024;
  s.total = off + 1024;
  return s;
}

// element (row r, column w) of a gy row held as four SWIZZLE_128B boxes [c rows][32 w]
__device__ __forceinline__ uint32_t wgs_elem_addr(uint32_t row_base, uint32_t box_bytes, int r, int w) {
  return row_base + (uint32_t)(w >> 5) * box_bytes + (uint32_t)r * 128 + ((((uint32_t)(w & 31) >> 2) ^ ((uint32_t)r & 7)) << 4) +
         ((uint32_t)w & 3) * 4;
}

__global__ void __launch_bounds__(kWgsThreads, 1)
conv_wgrad_shift_kernel(const __grid_constant__ CUtensorMap tm_gy, const __grid_constant__ CUtensorMap tm_x, WgParams p, WgsSmem L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + L.bar_off);
  uint64_t* rawfull = bars;       // [2] raw gy rows
  uint64_t* rawempty = bars + 2;  // [2]
  uint64_t* shfull = bars + 4;    // [2] shifted-operand stages (one arrive per shifter warp)
  uint64_t* shempty = bars + 6;   // [2]
  uint64_t* xfull = bars + 8;     // [kWgRing]
  uint64_t* xempty = bars + 12;   // [kWgRing]
  uint64_t* done = bars + 16;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 17);
  uint32_t* inited_s = (uint32_t*)(bars + 18);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.c, W = p.W, d = p.d;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_gy);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&rawfull[i], 1);
      mbar_init(&rawempty[i], 2);  // both MMA issuers
      mbar_init(&shfull[i], kWgsShifters);
      mbar_init(&shempty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 2);  // both MMA issuers read every x row
    }
    mbar_init(done, 2);
    *inited_s = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  // everything a don't-care operand half can reach must at least be finite
  for (uint32_t i = threadIdx.x; i < (L.ring_off - L.raw_off) / 4; i += blockDim.x) ((float*)(smem + L.raw_off))[i] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const bool zeros = p.boundary == PDEQX_PAD_ZEROS;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The x ring is laid out [32-column chunk c][slot][C rows]: for a fixed chunk the four slots are contiguous, so two row
    // taps in neighbouring slots form ONE K-major B operand of N = 2C. Dropped (dirichlet) rows are loaded as zeros
    // (out-of-range coordinate): every unit issues the same MMAs.
    if (lane == 0) {
      uint32_t pe = 0, used = 0, rpe = 0, rused = 0;
      int rb = 0;
      for (int64_t item = first; item < p.items; item += stride) {
        const WgItem it = wg_item(p, item);
        int v_next = it.j0 - 1;  // next chain row of x to fetch
        for (int jj = 0; jj < it.count; ++jj) {
          const int j = it.j0 + jj;
          const int h = it.r0 + j * p.d;
          const int v_hi = (j + 2 <= it.j0 + it.count) ? j + 2 : it.j0 + it.count;
          for (; v_next <= v_hi; ++v_next) {
            const int row = conv_map(it.r0 + v_next * p.d, p.H, p.boundary);  // -1: a row of zeros
            const int s = (v_next + 1) & (kWgRing - 1);
            if (used & (1u << s)) {
              mbar_wait_backoff(&xempty[s], (pe >> s) & 1u);
              pe ^= 1u << s;
            }
            used |= 1u << s;
            uint8_t* dst = smem + L.ring_off + (size_t)s * L.box_bytes;
            mbar_expect_tx(&xfull[s], L.slot_bytes);
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * L.slot_bytes, &tm_x, &xfull[s], c * 32, row, it.b * C);
          }
          // the raw gy row of this unit
          if (rused & (1u << rb)) {
            mbar_wait_backoff(&rawempty[rb], (rpe >> rb) & 1u);
            rpe ^= 1u << rb;
          }
          rused |= 1u << rb;
          uint8_t* dst = smem + L.raw_off + (size_t)rb * 4 * L.box_bytes;
          mbar_expect_tx(&rawfull[rb], 4 * L.box_bytes);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * L.box_bytes, &tm_gy, &rawfull[rb], c * 32, h, it.b * C);
          rb ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (whole warp convergent, one elected lane issues) =====================
    // Per k-step and A group: one MMA of N = 2C over the two row taps whose ring slots are neighbours + one of N = C for the
    // third -- 4 MMAs instead of 6. Accumulator columns are [group][row tap][C], so either pairing -- taps (0,1) or (1,2) --
    // lands on consistent columns. The two A groups are issued by TWO warps (warp 1: the shifted copies, warp 2 -- free after
    // the TMEM allocation -- the centre tap straight from the raw row): with one issuing thread the kernel was bound by that
    // thread's instruction stream (16 MMAs + barrier polling per 32-column chunk), not by the tensor pipe or HBM.
    const int grp = warp - 1;
    const uint32_t idesc1 = idesc_tf32(128, C, false, false), idesc2 = idesc_tf32(128, 2 * C, false, false);
    const uint64_t adesc0 = smem_desc_sw128(smem_u32(smem + L.sh_off), 0, 1024);
    const uint64_t rdesc0 = smem_desc_sw128(smem_u32(smem + L.raw_off), 0, 1024);
    const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem + L.ring_off), 0, 1024);
    const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)adesc0, r_lo0 = (uint32_t)rdesc0, b_lo0 = (uint32_t)bdesc0;
    const uint32_t box16 = L.box_bytes >> 4, slot16 = L.slot_bytes >> 4;
    uint32_t pf = 0, spf = 0, rpf = 0, started = 0;
    int st = 0, rb = 0;
    for (int64_t item = first; item < p.items; item += stride) {
      const WgItem it = wg_item(p, item);
      for (int jj = 0; jj < it.count; ++jj) {
        const int j = it.j0 + jj;
        const bool last = jj == it.count - 1;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (jj == 0 || t == 2) {  // first use of chain row v = j - 1 + t
            const int s = (j + t) & (kWgRing - 1);
            mbar_wait(&xfull[s], (pf >> s) & 1u);
            pf ^= 1u << s;
          }
        }
        if (grp == 1) {  // the centre tap reads the raw row itself
          mbar_wait(&rawfull[rb], (rpf >> rb) & 1u);
          rpf ^= 1u << rb;
        }
        // taps t = 0, 1, 2 sit in slots s0, s0 + 1, s0 + 2 (mod 4): the pair that does not wrap around the ring is merged
        const int s0 = j & (kWgRing - 1);
        const int pair_t = s0 == kWgRing - 1 ? 1 : 0;      // first tap of the merged pair
        const int single_t = pair_t == 0 ? 2 : 0;
        const int pair_s = (j + pair_t) & (kWgRing - 1), single_s = (j + single_t) & (kWgRing - 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (grp == 0) {  // the shifted copies of this chunk are in the operand stage
            mbar_wait(&shfull[st], (spf >> st) & 1u);
            spf ^= 1u << st;
          }
          tc_fence_after();
          if (elect_one()) {
            // group 0: A = [shift tw=0 ; shift tw=2] (operand stage); group 1: A = [centre ; don't-care] (raw row)
            const uint32_t a_lo = grp == 0 ? a_lo0 + (uint32_t)st * 2 * box16 : r_lo0 + (uint32_t)(rb * 4 + c) * box16;
            const uint32_t bp = b_lo0 + (uint32_t)c * slot16 + (uint32_t)pair_s * box16;
            const uint32_t bs = b_lo0 + (uint32_t)c * slot16 + (uint32_t)single_s * box16;
            const uint32_t dp = tmem_base + (uint32_t)((3 * grp + pair_t) * C), ds = tmem_base + (uint32_t)((3 * grp + single_t) * C);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc = (c | k) ? 1u : started;
              mma_tf32_lohi(dp, a_lo + k * 2, a_hi, bp + k * 2, b_hi, idesc2, acc);
              mma_tf32_lohi(ds, a_lo + k * 2, a_hi, bs + k * 2, b_hi, idesc1, acc);
            }
            if (grp == 0) mma_commit(&shempty[st]);
            if (c == 3) {
              // raw row: group 1 has read it; group 0's commit stands for the shifters (they are done with a raw row before
              // its last chunk can be consumed) -- two arrivals free the buffer
              mma_commit(&rawempty[rb]);
#pragma unroll
              for (int t = 0; t < 3; ++t)
                if (t == 0 || last) mma_commit(&xempty[(j + t) & (kWgRing - 1)]);
            }
          }
          __syncwarp();
          st ^= 1;
        }
        rb ^= 1;
        started = 1;
      }
    }
    if (elect_one()) {
      if (grp == 0) *inited_s = started ? 7u : 0u;
      mma_commit(done);
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ===================== shifters: raw gy row -> [shift tw=0 | shift tw=2] operand chunks =====================
    //   A_0[o, w'] = gy[o, w' + d] (+ the halo columns that reach x column w' through the boundary map), A_2[o, w'] = gy[o, w' - d] (...)
    // -- exactly what conv_shift_rows_kernel writes to HBM for the kernel above. One thread = one 16-byte unit (4 columns) of one
    // channel row; interior units are two 128-bit loads + a funnel, the d columns next to each edge take the scalar path.
    const int tid = (warp - 3) * 32 + lane;
    const uint32_t raw0 = smem_u32(smem + L.raw_off), sh0 = smem_u32(smem + L.sh_off);
    uint32_t rpf = 0, spe = 0, sused = 0;
    int st = 0, rb = 0;
    const int dm = d & 3;
    for (int64_t item = first; item < p.items; item += stride) {
      const WgItem it = wg_item(p, item);
      for (int jj = 0; jj < it.count; ++jj) {
        mbar_wait(&rawfull[rb], (rpf >> rb) & 1u);
        rpf ^= 1u << rb;
        const uint32_t rbase = raw0 + (uint32_t)rb * 4 * L.box_bytes;
        for (int c = 0; c < 4; ++c) {
          if (sused & (1u << st)) {
            mbar_wait(&shempty[st], (spe >> st) & 1u);
            spe ^= 1u << st;
          }
          sused |= 1u << st;
          const uint32_t sbase = sh0 + (uint32_t)st * 2 * L.box_bytes;
          // units of this chunk: 2 taps x C rows x 8 sixteen-byte units
#ifndef PDEQX_WG_DIAG  // (timing experiment: no shifter work)
          for (int u = tid; u < 2 * C * 8; u += kWgsShifters * 32) {
            const int tap = u >> 9;  // (C = 64) 0: tw = 0 (source column w' + d), 1: tw = 2 (source column w' - d); warp-uniform
            const int r = (u >> 3) & 63, cu = u & 7;
            const int w0 = c * 32 + cu * 4;  // first of the unit's four columns w'
            const int sh = tap == 0 ? d : -d;
            float o4[4];
            const int src0 = w0 + sh;
            // periodic: the operand is the circular shift of the row (direct part + wrapped halo), every unit takes the fast path;
            // dirichlet / neumann: units whose sources leave the row or that receive mirrored halo columns take the scalar path
            const bool wrap = p.boundary == PDEQX_PAD_WRAP;
            const bool interior = wrap || (src0 >= 0 && src0 + 3 < W && (zeros || (w0 > d && w0 + 3 < W - 1 - d)));
            if (interior) {
              // source columns src0 .. src0 + 3 = units ub and ub + 1 of the row's 32, offset dm (tap 0) or 4 - dm (tap 2)
              const int off = tap == 0 ? dm : ((4 - dm) & 3);
              const int ub = ((src0 - off) >> 2) & 31;
              float lo[4], hi[4];
              ld_shared_v4(wgs_elem_addr(rbase, L.box_bytes, r, ub * 4), lo);
              if (off) {
                ld_shared_v4(wgs_elem_addr(rbase, L.box_bytes, r, ((ub + 1) & 31) * 4), hi);
                if (off == 1) { o4[0] = lo[1]; o4[1] = lo[2]; o4[2] = lo[3]; o4[3] = hi[0]; }
                else if (off == 2) { o4[0] = lo[2]; o4[1] = lo[3]; o4[2] = hi[0]; o4[3] = hi[1]; }
                else { o4[0] = lo[3]; o4[1] = hi[0]; o4[2] = hi[1]; o4[3] = hi[2]; }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) o4[i] = lo[i];
              }
            } else {
              // zero-extended direct source + (neumann) the one mirrored halo column that maps onto x column w:
              //   tap 0: conv_map(e - d) = d - e = w  ->  gy[d - w] for 1 <= w <= d;   tap 2: conv_map(W + e) = W - 2 - e = w  ->
              //   gy[2W - 2 - d - w] for W - 1 - d <= w <= W - 2   (what conv_shift_rows_kernel's search loop finds)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int w = w0 + i, src = w + sh;
                float a = (src >= 0 && src < W) ? ld_shared_f32(wgs_elem_addr(rbase, L.box_bytes, r, src)) : 0.f;
                if (!zeros) {
                  const int hs = tap == 0 ? d - w : 2 * W - 2 - d - w;
                  const bool has = tap == 0 ? (w >= 1 && w <= d) : (w >= W - 1 - d && w <= W - 2);
                  if (has) a += ld_shared_f32(wgs_elem_addr(rbase, L.box_bytes, r, hs));
                }
                o4[i] = a;
              }
            }
            st_shared_v4(sbase + (uint32_t)tap * L.box_bytes + (uint32_t)r * 128 + (((uint32_t)cu ^ ((uint32_t)r & 7)) << 4), o4);
          }
#endif
          fence_proxy_async();  // the tensor core reads what this thread wrote through the async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(&shfull[st]);
          st ^= 1;
        }
        rb ^= 1;
      }
    }
    if (warp >= 4 && warp < 8) {
      // ===================== flush: TMEM -> scratch[cta][tw][th][i][o] =====================
      const int q = warp & 3;
      mbar_wait(done, 0);
      tc_fence_after();
      const uint32_t inited = *(volatile uint32_t*)inited_s;
      float* out = p.scratch + (size_t)blockIdx.x * 9 * C * C;
      const int o = (q & 1) * 32 + lane;
      for (int g = 0; g < 2; ++g) {
        if (g == 1 && q >= 2) break;
        const int tw = g == 0 ? (q < 2 ? 0 : 2) : 1;
        for (int t = 0; t < 3; ++t) {
          for (int c0 = 0; c0 < C; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g * 3 + t) * C + c0), v);
            tmem_ld_wait();
            if (o < C) {
#pragma unroll
              for (int jx = 0; jx < 16; ++jx)
                out[((size_t)(tw * 3 + t) * C + c0 + jx) * C + o] = ((inited >> t) & 1u) ? v[jx] : 0.f;
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

// ------------------------------------------------------------------------------------------
// Filter gradient with the column-shifted operands in TENSOR MEMORY (tcgen05.mma, A operand from TMEM).
// ------------------------------------------------------------------------------------------
// The kernel above is bound by the shared-memory operand fetch of its MMAs: every tcgen05.mma re-reads its 128-row A tile
// (4 KB per k-step) next to B, 28 KB per k-step over the four MMAs, and the shifter warps write 64 KB of shifted copies per
// row through the same pipe. Here the shifted copies never exist in shared memory: the shifter warps read the raw gy row
// (ld.shared.v4), shift in REGISTERS (a compile-time register offset per dilation residue) and write the A tile straight
// into tensor memory (tcgen05.st, lane = (column tap, o), column = w); the MMAs take A from TMEM and fetch only B -- the x
// rows of the ring, 6 KB per k-step and group -- so an M = 128, N = 128 + 64 pair runs at its 96-cycle math floor.
//   TMEM: columns 0..383 the six accumulators [group][row tap][C] as before; 384..447 A of group 0 ([shift tw=0 ; shift
//   tw=2], two 32-column stages), 448..511 A of group 1 ([centre ; zeros], two stages).
//   Shifter warps 4..15 = (role r, lane quarter q): r = 0 / 1 build group 0 for the even / odd 32-column chunks of a row,
//   r = 2 (q < 2) copies the centre tap and, having every gy element of its channel in registers once, sums the BIAS
//   gradient on the way (per-CTA partial sums, reduced with the filter-gradient slabs: no separate pass over gy).
//   The 32 KB of shifted-operand stages the other kernel keeps in shared memory hold a third raw gy row in flight here.
struct WgtSmem {
  uint32_t raw_off, ring_off, slot_bytes, box_bytes, bar_off, total;
};
constexpr int kWgtThreads = 512;
constexpr int kWgtACol = 384;   // first TMEM column of the A stages

static WgtSmem wgt_smem(int c, int ring, int raw) {
  WgtSmem s{};
  uint32_t off = 0;
  s.box_bytes = (uint32_t)c * 128;  // [c rows][32 w]
  s.raw_off = off;
  off += (uint32_t)raw * 4 * s.box_bytes;
  s.ring_off = off;
  s.slot_bytes = (uint32_t)ring * s.box_bytes;  // stride between the 32-column chunks of the ring: [chunk][slot][c rows]
  off += 4 * s.slot_bytes;
  s.bar_off = off;
  off += 1024;
  s.total = off + 1024;
  return s;
}

// RING = x rows resident (3 in use + RING - 3 in flight), RAW = raw gy rows in flight: (4, 3) or (5, 2) fill the 227 KB
template <int RING, int RAW>
__global__ void __launch_bounds__(kWgtThreads, 1)
conv_wgrad_tmem_kernel(const __grid_constant__ CUtensorMap tm_gy, const __grid_constant__ CUtensorMap tm_x, WgParams p, WgtSmem L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + L.bar_off);
  uint64_t* rawfull = bars;       // [RAW] raw gy rows
  uint64_t* rawempty = bars + 4;  // [RAW] (one arrive per reading shifter warp)
  uint64_t* afull = bars + 8;     // [2 groups][2 stages] A chunk is in TMEM
  uint64_t* aempty = bars + 12;   // [2 groups][2 stages] the MMAs that read it have completed
  uint64_t* xfull = bars + 16;    // [RING]
  uint64_t* xempty = bars + 24;   // [RING]
  uint64_t* done = bars + 32;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 33);
  uint32_t* inited_s = (uint32_t*)(bars + 34);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.c, W = p.W, d = p.d;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_gy);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < RAW; ++i) {
      mbar_init(&rawfull[i], 1);
      mbar_init(&rawempty[i], 10);  // 4 + 4 group-0 warps, 2 centre warps
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], 4);      // group 0: the four lane-quarter warps of the stage's role
      mbar_init(&afull[2 + i], 2);  // group 1: lanes 0..63
      mbar_init(&aempty[i], 1);
      mbar_init(&aempty[2 + i], 1);
    }
    for (int i = 0; i < RING; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 2);  // both MMA issuers read every x row
    }
    mbar_init(done, 2);
    *inited_s = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 14 || warp == 15) {
    // the don't-care upper half (lanes 64..127) of group 1's A stages: zeros once, so that nothing undefined is multiplied
    float z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0.f;
    tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kWgtACol + 64), z);
    tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kWgtACol + 96), z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const bool zeros = p.boundary == PDEQX_PAD_ZEROS;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // x ring laid out [32-column chunk c][slot][C rows] as in the kernel above (two row taps in neighbouring slots are one
    // K-major B operand of N = 2C); dropped (dirichlet) rows are loaded as zeros.
    if (lane == 0) {
      uint32_t pe = 0, used = 0;
      for (int64_t item = first; item < p.items; item += stride) {
        const WgItem it = wg_item(p, item);
        int v_next = it.j0 - 1;  // next chain row of x to fetch
        for (int jj = 0; jj < it.count; ++jj) {
          const int j = it.j0 + jj;
          const int v_hi = (j + RING - 2 <= it.j0 + it.count) ? j + RING - 2 : it.j0 + it.count;
          for (; v_next <= v_hi; ++v_next) {
            const int row = conv_map(it.r0 + v_next * p.d, p.H, p.boundary);  // -1: a row of zeros
            const int s = (v_next + 1) % RING;
            if (used & (1u << s)) {
              mbar_wait_backoff(&xempty[s], (pe >> s) & 1u);
              pe ^= 1u << s;
            }
            used |= 1u << s;
            uint8_t* dst = smem + L.ring_off + (size_t)s * L.box_bytes;
            mbar_expect_tx(&xfull[s], 4 * L.box_bytes);
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * L.slot_bytes, &tm_x, &xfull[s], c * 32, row, it.b * C);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== TMA producer of the raw gy rows =====================
    // A thread of its own: the x producer above blocks on a ring slot that the MMAs of the previous unit still read, and a gy
    // row requested only behind that wait arrives one full DRAM latency late in EVERY unit (ncu: the shifter warps' wait for
    // the raw row was the top stall, the tensor pipe 50 % idle). This thread waits for nothing but a free raw buffer, so
    // RAW rows are in flight.
    if (lane == 0) {
      uint32_t rpe = 0, rused = 0;
      int rb = 0;
      for (int64_t item = first; item < p.items; item += stride) {
        const WgItem it = wg_item(p, item);
        for (int jj = 0; jj < it.count; ++jj) {
          const int h = it.r0 + (it.j0 + jj) * p.d;
          if (rused & (1u << rb)) {
            mbar_wait_backoff(&rawempty[rb], (rpe >> rb) & 1u);
            rpe ^= 1u << rb;
          }
          rused |= 1u << rb;
          uint8_t* dst = smem + L.raw_off + (size_t)rb * 4 * L.box_bytes;
          mbar_expect_tx(&rawfull[rb], 4 * L.box_bytes);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * L.box_bytes, &tm_gy, &rawfull[rb], c * 32, h, it.b * C);
          rb = rb == RAW - 1 ? 0 : rb + 1;
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (whole warp convergent, one elected lane issues) =====================
    // warp 1: group 0 (A = [shift tw=0 ; shift tw=2]), warp 2: group 1 (A = [centre ; zeros]). Per 32-column chunk and group:
    // 4 k-steps x (one MMA of N = 2C over the two row taps in neighbouring ring slots + one of N = C for the third).
    const int grp = warp - 1;
    const uint32_t idesc1 = idesc_tf32(128, C, false, false), idesc2 = idesc_tf32(128, 2 * C, false, false);
    const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem + L.ring_off), 0, 1024);
    const uint32_t b_hi = (uint32_t)(bdesc0 >> 32), b_lo0 = (uint32_t)bdesc0;
    const uint32_t box16 = L.box_bytes >> 4, slot16 = L.slot_bytes >> 4;
    const uint32_t a_col0 = tmem_base + (uint32_t)(kWgtACol + grp * 64);
    uint32_t pf = 0, apf = 0, started = 0;
    int st = 0;
    for (int64_t item = first; item < p.items; item += stride) {
      const WgItem it = wg_item(p, item);
      for (int jj = 0; jj < it.count; ++jj) {
        const int j = it.j0 + jj;
        const bool last = jj == it.count - 1;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (jj == 0 || t == 2) {  // first use of chain row v = j - 1 + t
            const int s = (j + t) % RING;
            mbar_wait(&xfull[s], (pf >> s) & 1u);
            pf ^= 1u << s;
          }
        }
        // taps t = 0, 1, 2 sit in slots s0, s0 + 1, s0 + 2 (mod 4): the pair that does not wrap around the ring is merged
        const int s0 = j % RING;
        const int pair_t = s0 == RING - 1 ? 1 : 0;  // first tap of the merged pair
        const int single_t = pair_t == 0 ? 2 : 0;
        const int pair_s = (j + pair_t) % RING, single_s = (j + single_t) % RING;
        const uint32_t dp = tmem_base + (uint32_t)((3 * grp + pair_t) * C), ds = tmem_base + (uint32_t)((3 * grp + single_t) * C);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mbar_wait(&afull[grp * 2 + st], (apf >> st) & 1u);
          apf ^= 1u << st;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_t = a_col0 + (uint32_t)(st * 32);
            const uint32_t bp = b_lo0 + (uint32_t)c * slot16 + (uint32_t)pair_s * box16;
            const uint32_t bs = b_lo0 + (uint32_t)c * slot16 + (uint32_t)single_s * box16;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc = (c | k) ? 1u : started;
              mma_tf32_ts(dp, a_t + k * 8, bp + k * 2, b_hi, idesc2, acc);
              mma_tf32_ts(ds, a_t + k * 8, bs + k * 2, b_hi, idesc1, acc);
            }
            mma_commit(&aempty[grp * 2 + st]);
            if (c == 3) {
#pragma unroll
              for (int t = 0; t < 3; ++t)
                if (t == 0 || last) mma_commit(&xempty[(j + t) % RING]);
            }
          }
          __syncwarp();
          st ^= 1;
        }
        started = 1;
      }
    }
    if (elect_one()) {
      if (grp == 0) *inited_s = started ? 7u : 0u;
      mma_commit(done);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== shifters: raw gy row (shared memory) -> A chunks (tensor memory) =====================
    //   group 0, lane (tap, o):  tap 0 (tw = 0): A[o, w'] = gy[o, w' + d];  tap 1 (tw = 2): A[o, w'] = gy[o, w' - d]
    //   (+ the wrapped / mirrored halo columns that reach x column w' through the boundary map, as in the kernel above).
    // One thread = one channel row of one 32-column chunk: nine 128-bit loads cover the 32 source columns, which sit at a
    // warp-uniform register offset 0..3; the chunks whose sources leave the row (dirichlet / neumann edges) go element-wise.
    const int q = warp & 3, r = (warp - 4) >> 2;
    const uint32_t raw0 = smem_u32(smem + L.raw_off);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool wrap = p.boundary == PDEQX_PAD_WRAP;
    const int tap = q >> 1;
    const int o = (q & 1) * 32 + lane;  // group 0: channel of this lane; group 1 (q < 2): q * 32 + lane, the same number
    const int dm = d & 3;
    const int sh = tap == 0 ? d : -d;
    const int off = tap == 0 ? dm : ((4 - dm) & 3);
    uint32_t rpf = 0, eph = 0;
    int rb = 0;
    float bsum = 0.f, msum = 0.f;
    if (r < 2 || q < 2) {
      for (int64_t item = first; item < p.items; item += stride) {
        const WgItem it = wg_item(p, item);
        float rstd = 1.f;  // GroupNorm fold: the sample's 1/sigma
        if (p.gn_stats) rstd = __ldg(p.gn_stats + 2 * it.b + 1);
        for (int jj = 0; jj < it.count; ++jj) {
          mbar_wait(&rawfull[rb], (rpf >> rb) & 1u);
          rpf ^= 1u << rb;
          const uint32_t rbase = raw0 + (uint32_t)rb * 4 * L.box_bytes;
          if (r < 2) {
            // group 0, chunks c = r and r + 2: always stage r
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
              const int c = 2 * cc + r;
              const int w0 = c * 32, src0 = w0 + sh;
              float in[36];
              const bool interior = wrap || (src0 >= 0 && src0 + 31 < W && (zeros || (w0 > d && w0 + 31 < W - 1 - d)));
              if (interior) {
                const int ub = ((src0 - off) >> 2) & 31;  // first of the nine 16-byte units (circular in the row)
#pragma unroll
                for (int u = 0; u < 9; ++u) ld_shared_v4(wgs_elem_addr(rbase, L.box_bytes, o, ((ub + u) & 31) * 4), in + 4 * u);
                if (off == 1) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) in[i] = in[i + 1];
                } else if (off == 2) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) in[i] = in[i + 2];
                } else if (off == 3) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) in[i] = in[i + 3];
                }
              } else {
                // zero-extended direct source + (neumann) the one mirrored halo column that maps onto x column w:
                //   tap 0: conv_map(e - d) = d - e = w  ->  gy[d - w] for 1 <= w <= d;   tap 2: conv_map(W + e) = W - 2 - e = w  ->
                //   gy[2W - 2 - d - w] for W - 1 - d <= w <= W - 2
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const int w = w0 + i, src = w + sh;
                  float a = (src >= 0 && src < W) ? ld_shared_f32(wgs_elem_addr(rbase, L.box_bytes, o, src)) : 0.f;
                  if (!zeros) {
                    const int hs = tap == 0 ? d - w : 2 * W - 2 - d - w;
                    const bool has = tap == 0 ? (w >= 1 && w <= d) : (w >= W - 1 - d && w <= W - 2);
                    if (has) a += ld_shared_f32(wgs_elem_addr(rbase, L.box_bytes, o, hs));
                  }
                  in[i] = a;
                }
              }
              if (p.gn_stats) {  // rounded, not left to the tensor core's truncation: exactly the values the centre warps sum
#pragma unroll
                for (int i = 0; i < 32; ++i) in[i] = round_tf32(in[i] * rstd);
              }
              mbar_wait(&aempty[r], (eph & 1u) ^ 1u);  // the MMAs on this stage's previous chunk have completed
              eph ^= 1u;
              tc_fence_after();
              tmem_st32(lane_base + (uint32_t)(kWgtACol + r * 32), in);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&afull[r]);
            }
          } else {
            // group 1: the centre tap is the raw row itself; its running sum is the bias gradient
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              const int st = c & 1;
              float in[32];
#pragma unroll
              for (int u = 0; u < 8; ++u) ld_shared_v4(wgs_elem_addr(rbase, L.box_bytes, o, c * 32 + u * 4), in + 4 * u);
              float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int i = 0; i < 32; ++i) s4[i & 3] += in[i];
              const float rowsum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
              bsum += rowsum;
              if (p.gn_stats) {
                // M must cancel the mean part of what the tensor core accumulates EXACTLY: sum the operand values themselves
                float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  in[i] = round_tf32(in[i] * rstd);
                  r4[i & 3] += in[i];
                }
                msum = fmaf(__ldg(p.gn_stats + 2 * it.b), (r4[0] + r4[1]) + (r4[2] + r4[3]), msum);
              }
              mbar_wait(&aempty[2 + st], ((eph >> st) & 1u) ^ 1u);
              eph ^= 1u << st;
              tc_fence_after();
              tmem_st32(lane_base + (uint32_t)(kWgtACol + 64 + st * 32), in);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&afull[2 + st]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&rawempty[rb]);  // this warp's loads of the row have all landed in registers
          rb = rb == RAW - 1 ? 0 : rb + 1;
        }
      }
    }
    if (r == 2 && q < 2 && p.bias_part) p.bias_part[(size_t)blockIdx.x * C + o] = bsum;
    if (r == 2 && q < 2 && p.m_part) p.m_part[(size_t)blockIdx.x * C + o] = msum;
    if (r == 0) {
      // ===================== flush: TMEM -> scratch[cta][tw][th][i][o] =====================
      mbar_wait(done, 0);
      tc_fence_after();
      const uint32_t inited = *(volatile uint32_t*)inited_s;
      float* out = p.scratch + (size_t)blockIdx.x * 9 * C * C;
      for (int g = 0; g < 2; ++g) {
        if (g == 1 && q >= 2) break;
        const int tw = g == 0 ? (q < 2 ? 0 : 2) : 1;
        for (int t = 0; t < 3; ++t) {
          for (int c0 = 0; c0 < C; c0 += 16) {
            float v[16];
            tmem_ld16(lane_base + (uint32_t)((g * 3 + t) * C + c0), v);
            tmem_ld_wait();
            if (o < C) {
#pragma unroll
              for (int jx = 0; jx < 16; ++jx)
                out[((size_t)(tw * 3 + t) * C + c0 + jx) * C + o] = ((inited >> t) & 1u) ? v[jx] : 0.f;
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

// gw[o][i][th][tw] = sum_cta scratch[cta][tw][th][i][o]   (+ what the edge kernel adds afterwards);
// gb[o] = sum_cta bias_part[cta][o] when the filter-gradient kernel summed gy on the way (bias_part != nullptr)
// GroupNorm fold (m_part != nullptr; C = 64, blockDim = 256 so that a block covers all o for four i):
//   gw[o][i][t] = gamma_i acc + beta_i G[o] - gamma_i M[o],  G[o] = sum gy (the bias gradient), M[o] = sum mu_b rstd_b gy
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ gw, int nslabs, int C,
                                         const float* __restrict__ bias_part, float* __restrict__ gb,
                                         const float* __restrict__ m_part, const float* __restrict__ gamma,
                                         const float* __restrict__ beta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over [tw][th][i][o]
  const int total = 9 * C * C;
  __shared__ float gm_s[2][64];
  if (m_part) {  // (block-uniform) every block sums G and M itself: nslabs coalesced 256-byte rows each
    if (threadIdx.x < 128) {
      const int o = threadIdx.x & 63;
      const float* src = (threadIdx.x < 64 ? bias_part : m_part) + o;
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      int s = 0;
      for (; s + 4 <= nslabs; s += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) part[j] += __ldg(src + (size_t)(s + j) * C);
      }
      for (; s < nslabs; ++s) part[0] += __ldg(src + (size_t)s * C);
      gm_s[threadIdx.x >> 6][o] = (part[0] + part[1]) + (part[2] + part[3]);
    }
    __syncthreads();
  }
  if (idx >= total) {
    if (bias_part && idx < total + C) {
      const int o = idx - total;
      float acc = 0.f;
      for (int s = 0; s < nslabs; ++s) acc += __ldg(bias_part + (size_t)s * C + o);
      gb[o] = acc;
    }
    return;
  }
  // eight independent partial sums: the 148 slab reads of a thread are otherwise one latency-bound dependent chain
  float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int s = 0;
  for (; s + 8 <= nslabs; s += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) part[j] += __ldg(scratch + (size_t)(s + j) * total + idx);
  }
  for (; s < nslabs; ++s) part[0] += __ldg(scratch + (size_t)s * total + idx);
  const float acc = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
  const int o = idx % C;
  int r = idx / C;
  const int i = r % C;
  r /= C;
  const int th = r % 3, tw = r / 3;
  float v = acc;
  if (m_part) {
    const float ga = gamma ? __ldg(gamma + i) : 1.f, be = beta ? __ldg(beta + i) : 0.f;
    v = ga * (acc - gm_s[1][o]) + be * gm_s[0][o];
  }
  gw[((o * C + i) * 3 + th) * 3 + tw] = v;
}

// gb[o] = sum_{b,h,w} gy[b,o,h,w]: one block per (b, o) plane chunk, atomics on C addresses
__global__ void __launch_bounds__(256)
conv_bias_grad_kernel(const float* __restrict__ gy, float* __restrict__ gb, int C, int64_t n) {
  const int64_t plane = blockIdx.x;  // b * C + o
  const float4* src = (const float4*)(gy + plane * n);
  float acc = 0.f;
  for (int64_t v = threadIdx.x; v < n / 4; v += blockDim.x) {
    const float4 t = __ldg(src + v);
    acc += (t.x + t.y) + (t.z + t.w);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(gb + (int)(plane % C), t);
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// rows per work item: a divisor of H (so that no item has a ragged tail), as close to 32 as possible
static int conv_segment(int H) {
  if (H <= 64) return H;
  for (int s = 64; s >= 8; --s)
    if (H % s == 0 && s <= 32) return s;
  for (int s = 64; s >= 8; --s)
    if (H % s == 0) return s;
  return 0;
}

static bool tc_conv_shape_ok(const ConvGeom& g) {
  if (g.ndim != 2 || g.groups != 1) return false;
  if (g.stride[0] != 1 || g.stride[1] != 1) return false;
  if (g.kernel[0] != g.kernel[1] || g.kernel[0] != 3) return false;
  if (g.dilation[0] != g.dilation[1]) return false;
  const int d = g.dilation[0], pd = d * (g.kernel[0] - 1) / 2;
  for (int a = 0; a < 2; ++a)
    if (g.pad_lo[a] != pd || g.pad_hi[a] != pd) return false;
  if (g.spatial[1] != 128) return false;
  if (pd >= g.spatial[0] || d > tc::kMaxDil) return false;
  if ((g.c_in != 32 && g.c_in != 64) || g.c_out != 64) return false;
  if (g.c_in != g.c_out) return false;  // the data gradient swaps the roles; keep both directions on the fast path
  if (g.boundary == PDEQX_PAD_EDGE) return false;
  if (3 * (g.c_out / 2) * tc::kAcc > tc::kConvTmemCols) return false;
  return true;
}

// Caller-provided scratch of the tcgen05 passes (pdeqx_conv_workspace_bytes): [re-laid-out filter][partial filter gradients]
constexpr size_t kWprepFloats = 64 * 64 * 9;
constexpr size_t kWgSlabs = 160;  // one partial filter gradient [3 tw][3 th][64 i][64 o] per CTA of the filter-gradient kernel
size_t tc_conv_workspace_bytes(const ConvGeom& g) {
  if (!tc_conv_available(g)) return 0;
  return sizeof(float) * (kWprepFloats + kWgSlabs * (9 * 64 * 64 + 2 * 64) + (size_t)g.batch * 64) + 256;
}
static float* ws_wprep(void* ws) { return (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255); }
static float* ws_wgrad(void* ws) { return ws_wprep(ws) + kWprepFloats; }
static float* ws_bias_part(void* ws) { return ws_wgrad(ws) + kWgSlabs * 9 * 64 * 64; }  // [slab][64] per-CTA sums of gy
static float* ws_m_part(void* ws) { return ws_bias_part(ws) + kWgSlabs * 64; }  // [slab][64] per-CTA sums of mu_b rstd_b gy (GroupNorm fold)
static float* ws_bias_b(void* ws) { return ws_m_part(ws) + kWgSlabs * 64; }     // [batch][64] per-sample epilogue constants (GroupNorm fold)

bool tc_conv_available(const ConvGeom& g) {
  if (!g_fast_paths.load(std::memory_order_relaxed)) return false;
  if (g.precision != PDEQX_PREC_TF32) return false;
  if (!tc_conv_shape_ok(g)) return false;
  if (!pdeqx_has_tcgen05() || !tc::encode_tiled_fn()) return false;
  return true;
}

static std::atomic<int64_t> g_pair_launches{0};  // diagnostic counter
// devices (bit = ordinal) that rejected the two-CTA cluster launch once -- a cached capability, like the attribute flags
static std::atomic<uint64_t> g_pair_unavailable{0};
int64_t tc_conv_pair_launches() { return g_pair_launches.load(std::memory_order_relaxed); }

static int check_conv_ws(const ConvGeom& g, void* ws, size_t ws_bytes) {
  if (!ws || ws_bytes < tc_conv_workspace_bytes(g)) {
    set_error("conv workspace too small: need %zu bytes (pdeqx_conv_workspace_bytes), got %zu", tc_conv_workspace_bytes(g), ws ? ws_bytes : (size_t)0);
    return PDEQX_ERR_WORKSPACE;
  }
  return PDEQX_OK;
}

// GroupNorm fold of the pair kernel (mode 1: forward, mode 2: data gradient; see conv_rows_pair_kernel)
struct GnFold {
  int mode = 0;
  const float* stats = nullptr;  // [B][2] mean, rstd of the norm's input
  const float* gamma = nullptr;  // [C] (NULL: ones)
  const float* beta = nullptr;   // [C] (NULL: zeros)
  const float* h = nullptr;      // mode 2: the norm's input
  float* parts = nullptr;        // partial sums out, [B][slots][2]
};

// The CTA-pair kernel's work split: chains of equal length cut into equal segments (both CTAs of a pair issue the same
// schedule) and an even number of items. Returns the segment length (0: the shape does not take the pair kernel).
static int conv_pair_segment(const ConvGeom& g, int c_in, int c_out, int sms) {
  const int H = g.spatial[0], d = g.dilation[0];
  if (!(c_in == 64 && c_out == 64 && H % d == 0 && sms >= 2)) return 0;
  const int chain = H / d, pairs = sms / 2;
  int best_seg = 0;
  int64_t best_cost = 0;
  for (int seg = 4; seg <= 64 && seg <= chain; ++seg) {
    if (chain % seg) continue;
    const int64_t items2 = (int64_t)g.batch * d * (chain / seg);
    if (items2 & 1) continue;
    const int64_t rounds = ceil_div(items2 / 2, (int64_t)pairs);
    const int64_t cost = rounds * (seg + 2);  // rows a CTA walks, the two halo rows of every segment included
    if (!best_seg || cost < best_cost) { best_seg = seg; best_cost = cost; }
  }
  return best_seg;
}

static int conv_rows_launch(const ConvGeom& g, const float* x, const float* w, bool flip, const float* bias,
                            const float* residual, int act, float* y, void* ws, size_t ws_bytes, cudaStream_t s,
                            const float* mask = nullptr, const GnFold* gn = nullptr) {
  int dev = 0, sms = 148;
  PDEQX_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PDEQX_TRY(check_conv_ws(g, ws, ws_bytes));
  float* wprep = ws_wprep(ws);
  const uint64_t dev_bit = 1ull << (dev & 63);
  const int c_in = flip ? g.c_out : g.c_in, c_out = flip ? g.c_in : g.c_out;
  const int k = g.kernel[0];
  PDEQX_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && (!residual || ((uintptr_t)residual & 15) == 0),
                  "tc_conv: tensors must be 16-byte aligned");
  const int total_w = c_out * c_in * k * k;
  const int gn_mode = gn ? gn->mode : 0;
  tc::conv_prep_weights_kernel<<<(total_w + 255) / 256, 256, 0, s>>>(w, wprep, c_out, c_in, k, flip ? 1 : 0,
                                                                     gn_mode == 1 ? gn->gamma : nullptr);
  PDEQX_LAUNCHED();
  if (gn_mode == 1) {
    tc::conv_gn_bias_kernel<<<c_out, 128, 0, s>>>(w, bias, gn->gamma, gn->beta, gn->stats, ws_bias_b(ws), c_out, c_in, k * k, g.batch);
    PDEQX_LAUNCHED();
  }
  const int H = g.spatial[0], W = g.spatial[1];
  CUtensorMap tm_x, tm_w, tm_y, tm_res;
  {
    uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)g.batch * c_in};
    uint64_t str[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
    uint32_t box[3] = {32, 1, (uint32_t)c_in};
    PDEQX_TRY(tc::make_tensor_map(&tm_x, x, 3, dims, str, box, tc::TM_SW128_ATOM32));
  }
  {
    const uint32_t N = (uint32_t)k * (c_out / 2);
    uint64_t dims[2] = {(uint64_t)c_in, (uint64_t)2 * k * N};
    uint64_t str[1] = {(uint64_t)c_in * 4};
    uint32_t box[2] = {32, N};
    PDEQX_TRY(tc::make_tensor_map(&tm_w, wprep, 2, dims, str, box, tc::TM_SW128));
  }
  {
    uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)g.batch * c_out};
    uint64_t str[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
    uint32_t box[3] = {32, 1, (uint32_t)tc::kConvOcnt};  // one epilogue warp's piece
    PDEQX_TRY(tc::make_tensor_map(&tm_y, y, 3, dims, str, box, tc::TM_SW128));
    if (residual) PDEQX_TRY(tc::make_tensor_map(&tm_res, residual, 3, dims, str, box, tc::TM_SW128));
    else tm_res = tm_y;
  }
  tc::ConvParams cp{};
  cp.H = H; cp.W = W; cp.c_in = c_in; cp.c_out = c_out; cp.k = k; cp.d = g.dilation[0]; cp.p = g.pad_lo[0];
  cp.boundary = g.boundary;
  cp.batch = g.batch;
  const int max_chain = (int)ceil_div(H, cp.d);
  cp.seg = max_chain < 32 ? max_chain : 32;
  cp.segs_per_image = (int)ceil_div(max_chain, cp.seg);  // segments per chain
  cp.act = act;
  cp.has_residual = residual ? 1 : 0;
  cp.has_bias = bias ? 1 : 0;
  cp.bias = bias;
  cp.y = y;
  cp.res = residual;
  cp.mask = mask;
  if (gn_mode == 1) {
    cp.gn_stats = gn->stats;
    cp.bias_b = ws_bias_b(ws);
    cp.has_bias = 0;
  } else if (gn_mode == 2) {
    cp.bias = gn->gamma;  // bias_s carries gamma (ones when the norm has no affine part: has_bias = 0 would zero it)
    cp.has_bias = gn->gamma ? 1 : 0;
    cp.mask = gn->h;
  }
  if (gn_mode) cp.parts = gn->parts;
  cp.items = (int64_t)g.batch * cp.d * cp.segs_per_image * 2;
  tc::ConvSmem L = tc::conv_smem(c_in, c_out, k);
  PDEQX_CHECK_ARG(L.total <= 227 * 1024, "tc_conv: shared memory budget exceeded (%u bytes)", L.total);
  // CTA-pair kernel (cta_group::2, all 64 output channels per unit): needs chains of equal length cut into equal segments
  // (both CTAs of a pair issue the same schedule) and an even number of items
  static const bool pair_enabled = [] { const char* e = getenv("PDEQX_CONV_PAIR"); return !(e && e[0] == '0'); }();
  const tc::ConvParams cp_one_cta = cp;
  if ((pair_enabled || gn_mode) && !(g_pair_unavailable.load(std::memory_order_relaxed) & dev_bit)) {
    const int chain = H / cp.d;
    const int best_seg = conv_pair_segment(g, c_in, c_out, sms);
    if (best_seg) {
      cp.seg = best_seg;
      cp.segs_per_image = chain / best_seg;
      cp.items = (int64_t)g.batch * cp.d * cp.segs_per_image;
      cp.slots = cp.d * cp.segs_per_image * 16;
      int grid2 = (int)(cp.items < sms ? cp.items : sms) & ~1;
      {  // an error pending from an earlier asynchronous failure must not be read as "cluster launch rejected" below
        const cudaError_t stale = cudaPeekAtLastError();
        if (stale != cudaSuccess) {
          set_error("CUDA error pending before the conv launch: %s", cudaGetErrorString(stale));
          return PDEQX_ERR_CUDA;
        }
      }
#define PDEQX_CONV_PAIR_LAUNCH(A, G)                                                                                      \
  do {                                                                                                                    \
    static DeviceOnce attr;                                                                                               \
    if (attr.first_use()) {                                                                                               \
      PDEQX_CUDA(cudaFuncSetAttribute(tc::conv_rows_pair_kernel<A, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    }                                                                                                                     \
    tc::conv_rows_pair_kernel<A, G><<<grid2, tc::kConvThreads, L.total, s>>>(tm_x, tm_w, tm_y, tm_res, cp, L);            \
  } while (0)
      if (gn_mode == 2) PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_IDENTITY, 2);
      else if (gn_mode == 1) {
        if (act == PDEQX_ACT_RELU) PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_RELU, 1);
        else if (act == PDEQX_ACT_GELU_TANH) PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_GELU_TANH, 1);
        else PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_IDENTITY, 1);
      } else if (act == PDEQX_ACT_RELU) PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_RELU, 0);
      else if (act == PDEQX_ACT_GELU_TANH) PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_GELU_TANH, 0);
      else PDEQX_CONV_PAIR_LAUNCH(PDEQX_ACT_IDENTITY, 0);
#undef PDEQX_CONV_PAIR_LAUNCH
      // A device that cannot co-schedule the two-CTA cluster with 227 KB each (a partitioned GPU, an odd SM count) rejects
      // the launch synchronously: remember that and take the one-CTA kernel from now on instead of failing the call.
      const cudaError_t pair_err = cudaGetLastError();  // (an error left over from an earlier call was surfaced by that call's own check)
      if (pair_err == cudaSuccess) {
        ::pdeqx::g_launches.fetch_add(1, std::memory_order_relaxed);
        g_pair_launches.fetch_add(1, std::memory_order_relaxed);
        return PDEQX_OK;
      }
      g_pair_unavailable.fetch_or(dev_bit, std::memory_order_relaxed);
      cp = cp_one_cta;
    }
  }
  if (gn_mode) {
    set_error("the GroupNorm fold needs the CTA-pair convolution kernel, which this shape / device does not take (pdeqx_conv_gn_fusable)");
    return PDEQX_ERR_UNSUPPORTED;
  }
  int grid = (int)(cp.items < sms ? cp.items : sms);
  grid &= ~1;  // even: a CTA keeps one channel half (its weights stay resident)
  PDEQX_CHECK_ARG(grid >= 2, "tc_conv: grid too small");
#define PDEQX_CONV_LAUNCH(A)                                                                                         \
  do {                                                                                                               \
    static DeviceOnce attr;                                                                                              \
    if (attr.first_use()) {                                                                                                     \
      PDEQX_CUDA(cudaFuncSetAttribute(tc::conv_rows_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                                                                                                                     \
    }                                                                                                                \
    tc::conv_rows_kernel<A><<<grid, tc::kConvThreads, L.total, s>>>(tm_x, tm_w, tm_y, tm_res, cp, L);                 \
  } while (0)
  if (act == PDEQX_ACT_RELU) PDEQX_CONV_LAUNCH(PDEQX_ACT_RELU);
  else if (act == PDEQX_ACT_GELU_TANH) PDEQX_CONV_LAUNCH(PDEQX_ACT_GELU_TANH);
  else PDEQX_CONV_LAUNCH(PDEQX_ACT_IDENTITY);
#undef PDEQX_CONV_LAUNCH
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

int tc_conv_fwd(const ConvGeom& g, const float* x, const float* w, const float* b, const float* residual, int act,
                float* y, void* ws, size_t ws_bytes, cudaStream_t s) {
  TimedScope timed(PDEQX_K_CONV_FWD, s);
  return conv_rows_launch(g, x, w, false, b, residual, act, y, ws, ws_bytes, s);
}

bool tc_conv_bwd_data_available(const ConvGeom& g) { return tc_conv_available(g); }

int tc_conv_bwd_data(const ConvGeom& g, const float* gy, const float* w, const float* add, const float* relu_mask, float* gx,
                     void* ws, size_t ws_bytes, cudaStream_t s) {
  TimedScope timed(PDEQX_K_CONV_BWD_DATA, s);
  if (g.boundary != PDEQX_PAD_REFLECT)
    return conv_rows_launch(g, gy, w, true, nullptr, add, PDEQX_ACT_IDENTITY, gx, ws, ws_bytes, s, relu_mask);
  // neumann: the adjoint of a reflect pad is a FOLD, not a convolution. The zero-extended part is the same implicit GEMM
  // (dirichlet data gradient); what arrives through mirrored halo positions only touches a frame of width d and is added
  // by the CUDA-core gather (the relu mask, if any, is applied by the caller after that).
  ConvGeom z = g;
  z.boundary = PDEQX_PAD_ZEROS;
  PDEQX_TRY(conv_rows_launch(z, gy, w, true, nullptr, add, PDEQX_ACT_IDENTITY, gx, ws, ws_bytes, s));
  return conv_bwd_data_mirror_part(g, gy, w, gx, s);
}

bool tc_conv_bwd_filter_available(const ConvGeom& g) { return tc_conv_available(g) && g.c_in == 64; }

// gw, gb zeroed by the caller (conv.cu). The column-shifted operands (halo folded in) are built in shared memory by the
// kernel itself: x and gy are read once. Every CTA flushes its partial sum to a slab of the workspace; a second kernel
// reduces the slabs (deterministic, no atomics on the filter gradient).
static int conv_bwd_filter_launch(const ConvGeom& g, const float* x, const float* gy, float* gw, float* gb, void* ws,
                                  size_t ws_bytes, cudaStream_t s, const GnFold* gn) {
  TimedScope timed(PDEQX_K_CONV_BWD_FILTER, s);
  int dev = 0, sms = 148;
  PDEQX_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PDEQX_CHECK_ARG(sms <= (int)kWgSlabs, "unexpected device (%d SMs)", sms);
  PDEQX_TRY(check_conv_ws(g, ws, ws_bytes));
  const int C = g.c_in, H = g.spatial[0], W = g.spatial[1], d = g.dilation[0];
  PDEQX_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)gy & 15) == 0, "tc_conv: tensors must be 16-byte aligned");
  CUtensorMap tm_gy, tm_x;
  uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)g.batch * C};
  uint64_t str[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
  uint32_t box[3] = {32, 1, (uint32_t)C};
  PDEQX_TRY(tc::make_tensor_map(&tm_gy, gy, 3, dims, str, box, tc::TM_SW128));
  PDEQX_TRY(tc::make_tensor_map(&tm_x, x, 3, dims, str, box, tc::TM_SW128));
  tc::WgParams wp{};
  wp.H = H; wp.W = W; wp.c = C; wp.d = d; wp.boundary = g.boundary;
  const int max_chain = (int)ceil_div(H, d);
  wp.seg = max_chain < 32 ? max_chain : 32;
  wp.segs_per_chain = (int)ceil_div(max_chain, wp.seg);
  wp.items = (int64_t)g.batch * d * wp.segs_per_chain;
  wp.scratch = ws_wgrad(ws);
  // Work items: chain segments of `seg` rows (+ one halo row at either end). A divisor of the chain length that balances
  // the CTAs: cost = rounds of the busiest CTA x (seg + 1).
  {
    const int chain = max_chain;
    int best = 0;
    int64_t best_cost = 0;
    for (int seg = 4; seg <= 64 && seg <= chain; ++seg) {
      if (chain % seg) continue;
      const int64_t items = (int64_t)g.batch * d * (chain / seg);
      const int64_t cost = ceil_div(items, (int64_t)sms) * (seg + 1);
      if (!best || cost < best_cost) { best = seg; best_cost = cost; }
    }
    if (best && H % d == 0) {
      wp.seg = best;
      wp.segs_per_chain = chain / best;
      wp.items = (int64_t)g.batch * d * wp.segs_per_chain;
    }
  }
  const int grid = (int)(wp.items < sms ? wp.items : sms);
  // A operand from tensor memory (default) / shifted copies in shared memory (PDEQX_WG_TMEM=0: the round-1 kernel)
  static const bool tmem_a = [] { const char* e = getenv("PDEQX_WG_TMEM"); return !(e && e[0] == '0'); }();
  if (tmem_a) {
    static const bool ring5 = [] { const char* e = getenv("PDEQX_WG_RING5"); return e && e[0] == '1'; }();
    tc::WgtSmem L = ring5 ? tc::wgt_smem(C, 5, 2) : tc::wgt_smem(C, 4, 3);
    PDEQX_CHECK_ARG(L.total <= 227 * 1024, "tc_conv_bwd_filter: shared memory budget exceeded (%u bytes)", L.total);
    wp.bias_part = (gb || gn) ? ws_bias_part(ws) : nullptr;
    if (gn) {
      wp.gn_stats = gn->stats;
      wp.m_part = ws_m_part(ws);
    }
    static DeviceOnce attr;
    if (attr.first_use()) {
      PDEQX_CUDA(cudaFuncSetAttribute(tc::conv_wgrad_tmem_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      PDEQX_CUDA(cudaFuncSetAttribute(tc::conv_wgrad_tmem_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (ring5) tc::conv_wgrad_tmem_kernel<5, 2><<<grid, tc::kWgtThreads, L.total, s>>>(tm_gy, tm_x, wp, L);
    else tc::conv_wgrad_tmem_kernel<4, 3><<<grid, tc::kWgtThreads, L.total, s>>>(tm_gy, tm_x, wp, L);
    PDEQX_LAUNCHED();
    tc::conv_wgrad_reduce_kernel<<<(9 * C * C + C + 255) / 256, 256, 0, s>>>(wp.scratch, gw, grid, C, gb ? wp.bias_part : nullptr, gb,
                                                                             wp.m_part, gn ? gn->gamma : nullptr, gn ? gn->beta : nullptr);
    PDEQX_LAUNCHED();
    return PDEQX_OK;
  }
  if (gn) {
    set_error("the GroupNorm fold needs the tensor-memory filter-gradient kernel (PDEQX_WG_TMEM=0 is set)");
    return PDEQX_ERR_UNSUPPORTED;
  }
  tc::WgsSmem L = tc::wgs_smem(C);
  PDEQX_CHECK_ARG(L.total <= 227 * 1024, "tc_conv_bwd_filter: shared memory budget exceeded (%u bytes)", L.total);
  static DeviceOnce attr;
  if (attr.first_use())
    PDEQX_CUDA(cudaFuncSetAttribute(tc::conv_wgrad_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  tc::conv_wgrad_shift_kernel<<<grid, tc::kWgsThreads, L.total, s>>>(tm_gy, tm_x, wp, L);
  PDEQX_LAUNCHED();
  tc::conv_wgrad_reduce_kernel<<<(9 * C * C + 255) / 256, 256, 0, s>>>(wp.scratch, gw, grid, C, nullptr, nullptr, nullptr, nullptr, nullptr);
  PDEQX_LAUNCHED();
  if (gb) {
    PDEQX_TRY(zero_small(gb, C, nullptr, 0, s));
    tc::conv_bias_grad_kernel<<<g.batch * C, 256, 0, s>>>(gy, gb, C, (int64_t)H * W);
    PDEQX_LAUNCHED();
  }
  return PDEQX_OK;
}

int tc_conv_bwd_filter(const ConvGeom& g, const float* x, const float* gy, float* gw, float* gb, void* ws, size_t ws_bytes,
                       cudaStream_t s) {
  return conv_bwd_filter_launch(g, x, gy, gw, gb, ws, ws_bytes, s, nullptr);
}

// ------------------------------------------------------------------------------------------
// GroupNorm (groups = 1) -> PhysicsConv -> activation with the norm folded into the convolution's three passes
// (pdequinox/blocks/_dilated_res_block.py:141-151). The normalised tensor is never written or read.
// ------------------------------------------------------------------------------------------
static int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

bool tc_conv_gn_fusable(const ConvGeom& g) {
  if (!tc_conv_available(g) || !tc_conv_bwd_filter_available(g)) return false;
  if (g.boundary != PDEQX_PAD_WRAP) return false;  // the norm's shift is a per-sample constant only if every tap reads a real pixel
  static const bool tmem_a = [] { const char* e = getenv("PDEQX_WG_TMEM"); return !(e && e[0] == '0'); }();
  if (!tmem_a) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (g_pair_unavailable.load(std::memory_order_relaxed) & (1ull << (dev & 63))) return false;
  return conv_pair_segment(g, g.c_in, g.c_out, device_sms()) != 0;
}

int tc_conv_gn_slots(const ConvGeom& g) {
  const int seg = conv_pair_segment(g, g.c_in, g.c_out, device_sms());
  if (!seg) return 0;
  return g.dilation[0] * (g.spatial[0] / g.dilation[0] / seg) * 16;
}

int tc_conv_gn_fwd(const ConvGeom& g, const float* x, const float* w, const float* b, const float* stats, const float* gamma,
                   const float* beta, int act, float* y, float* out_parts, void* ws, size_t ws_bytes, cudaStream_t s) {
  TimedScope timed(PDEQX_K_CONV_FWD, s);
  GnFold gn;
  gn.mode = 1; gn.stats = stats; gn.gamma = gamma; gn.beta = beta; gn.parts = out_parts;
  return conv_rows_launch(g, x, w, false, b, nullptr, act, y, ws, ws_bytes, s, nullptr, &gn);
}

int tc_conv_gn_bwd_filter(const ConvGeom& g, const float* x, const float* gy, const float* stats, const float* gamma,
                          const float* beta, float* gw, float* gb, void* ws, size_t ws_bytes, cudaStream_t s) {
  GnFold gn;
  gn.mode = 1; gn.stats = stats; gn.gamma = gamma; gn.beta = beta;
  return conv_bwd_filter_launch(g, x, gy, gw, gb, ws, ws_bytes, s, &gn);
}

int tc_conv_gn_bwd_data(const ConvGeom& g, const float* gy, const float* w, const float* gamma, const float* h, float* v,
                        float* parts, void* ws, size_t ws_bytes, cudaStream_t s) {
  TimedScope timed(PDEQX_K_CONV_BWD_DATA, s);
  GnFold gn;
  gn.mode = 2; gn.gamma = gamma; gn.h = h; gn.parts = parts;
  return conv_rows_launch(g, gy, w, true, nullptr, nullptr, PDEQX_ACT_IDENTITY, v, ws, ws_bytes, s, nullptr, &gn);
}

}  // namespace pdeqx

