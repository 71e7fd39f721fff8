// Per-mode complex channel mix of SpectralConv and its two adjoints on tcgen05 (TF32 precision).
//
//   replaces the einsum "i...,oi...->o..." over the 2^(D-1) corner blocks
//   (pdequinox/conv/_spectral_conv.py:129,133-135) and its two reverse passes.
//
// Per retained mode p the op is a small complex GEMM  C[m,n] = sum_k A(m,k) * B(n,k)^(conj)  (M, N, K = batch / channels).
// The reference keeps the MODE index contiguous in every operand (mode space [B, C, J] complex-interleaved, weights
// (G, o, i, *m) planar), so no operand has 16 contiguous bytes along K or M/N of a mode's matrix: TMA cannot deliver a UMMA
// tile. A CTA owns 8 consecutive modes (32-byte runs of a weight plane, 64-byte runs of mode space); one thread streams RAW
// chunks of 4 K indices with TMA (boxes with the mode axis innermost, zero-filled past the batch / channel counts, 3 - 5
// chunks in flight), and 16 warps TRANSPOSE them: 16-byte shared loads -> registers -> scalar stores into the K-major
// no-swizzle UMMA layout (8-row x 16-byte core matrices) of eight per-mode operand tiles. The complex product runs
// as ONE real MMA per mode and K step with the parts stacked along M and K:
//   A' row 2m   = [ Re A(m, k..k+3) | s1 Im A(m, k..k+3) ]      B' row n = [ Re B(n, k..k+3) | Im B(n, k..k+3) ]
//   A' row 2m+1 = [ Im A(m, k..k+3) | s2 Re A(m, k..k+3) ]      (s1, s2) = (-1, +1), conj(B): (+1, -1)
//   D = A' B'^T : row 2m = Re C(m, :), row 2m+1 = Im C(m, :)    (M = 128 = 64 complex rows, N = NT, K = 8 per chunk)
// so the eight accumulators (8 x NT TMEM columns) hold re / im of one output element in neighbouring lanes of one warp:
// the epilogue pairs them with one shuffle per value and stores whole 32-byte runs along the mode axis.
// The strides that the no-swizzle layout leaves free (SBO between 8-row groups, the gap between per-mode tiles) are
// padded so that every transposing warp store hits 32 distinct banks.
#include "tc_common.cuh"
#include "spectral_plan.h"
#include "modemap.cuh"

namespace pdeqx {
namespace tc {

struct MixArgs {
  const float2* a;     // interleaved: A(m,k,p) = a[m * a_sm + k * a_sk + p]
  int64_t a_sm, a_sk;  // (float2 units)
  const float2* b2;    // interleaved B (dW)
  const float* br;     // planar B (mix): B(n,k,p) = br/bi[n * b_sn + k * b_sk + woff(p)]
  const float* bi;
  int64_t b_sn, b_sk;  // float2 units (interleaved) / floats (planar)
  float2* c2;          // interleaved C (mix)
  float* cr;           // planar C (dW)
  float* ci;
  int64_t c_sm, c_sn;
  int M, N, K;
  int64_t w_corner;  // O * I * mprod: weight floats per corner
  uint32_t a_rm, a_rk, b_rn, b_rk;  // byte strides of (row, k) inside a raw chunk (set by the launcher)
  int a_k_inner, b_k_inner;         // k is the faster of the two box dimensions
  ModeMap mm;
};

constexpr int MX_PT = 8;        // modes per CTA
constexpr int MX_KC = 4;        // complex K per chunk = 8 real K = one MMA per mode
constexpr int MX_PROD = 512;    // transposer threads (16 warps; they also run the epilogue)
constexpr uint32_t MX_LBO = 128;           // bytes between the two core matrices of one K step
constexpr uint32_t MX_SBO = 2 * 128 + 16;  // bytes between 8-row groups (+16: row groups 4 banks apart)
constexpr uint32_t MX_TILE_A = 16 * MX_SBO + 16;  // one mode's A' tile (128 rows), +16: tiles 4 banks apart
template <int NT>
__host__ __device__ constexpr uint32_t mx_tile_b() { return (NT / 8) * MX_SBO + 16; }
template <int NT>
__host__ __device__ constexpr uint32_t mx_stage_bytes() { return MX_PT * (MX_TILE_A + mx_tile_b<NT>()); }

// shared-memory matrix descriptor, no swizzle (layout_type 0), K-major: core matrix = 8 rows x 16 B stored contiguously;
// lbo = bytes between core matrices adjacent in K, sbo = bytes between 8-row groups along M/N
__device__ __forceinline__ uint64_t smem_desc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  return d;
}

// A rows of a warp's pass: complex rows m and m + 4 (row groups 4 banks apart) instead of m and m + 1 (8 banks apart, where
// the next mode pair's tile already sits)
__device__ __forceinline__ int mx_row(int mi) { return ((mi & 1) << 2) + ((mi >> 1) & 3) + ((mi >> 3) << 3); }

// raw operand chunk as TMA leaves it: A 64 rows x 4 k x 64 B (8 modes complex), B NT rows x 4 k x 64 B (interleaved) or two
// planes of NT x 4 x 32 B (planar); which of (row, k) is the faster box dimension depends on the operand (byte strides
// a_rm / a_rk / b_rn / b_rk in MixArgs)
constexpr uint32_t MX_RAW_A = 64 * MX_KC * 64;
template <int NT>
__host__ __device__ constexpr uint32_t mx_raw_bytes() { return MX_RAW_A + NT * MX_KC * 64; }

template <int NT, int NS, int R, bool B_PLANAR, bool C_PLANAR, bool CONJ_B>
__global__ void __launch_bounds__(MX_PROD + 64, 1)
mode_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b0,
                    const __grid_constant__ CUtensorMap tm_b1, MixArgs t) {
  pdl_enter();
  constexpr uint32_t TILE_B = mx_tile_b<NT>();
  constexpr uint32_t STAGE = mx_stage_bytes<NT>();
  constexpr uint32_t B_OFF = MX_PT * MX_TILE_A;  // B' tiles follow the A' tiles inside a stage
  constexpr uint32_t RAW = mx_raw_bytes<NT>();
  constexpr uint32_t RAW_OFF = NS * STAGE;        // raw chunks follow the operand stages
  constexpr int TCOLS = MX_PT * NT;               // 256 or 512 accumulator columns
  constexpr int NBP = (NT * MX_KC * 2 + MX_PROD - 1) / MX_PROD;  // planar B: 16-byte pieces per thread and plane
  constexpr int NBI = NT * MX_KC * 4 / MX_PROD;   // interleaved B: 16-byte pieces per thread
  extern __shared__ __align__(1024) uint8_t mx_smem[];
  __shared__ uint64_t full[NS], empty[NS], rawfull[R], rawempty[R], done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t p0 = (int64_t)blockIdx.x * MX_PT;
  const int n0 = blockIdx.y * NT, m0 = blockIdx.z * 64;

  // validity of the tile's two 4-mode granules (the last-axis mode count is a multiple of 4) and the weight offset
  int g0, g1;
  int64_t loc0, loc1;
  const bool gv0 = mode_to_weight(t.mm, p0, g0, loc0);
  const bool gv1 = mode_to_weight(t.mm, p0 + 4, g1, loc1);
  const int64_t woff = (int64_t)g0 * t.w_corner + loc0;  // planar operands: offset of mode p0 inside an (n,k) plane
  if (!gv0) {
    // padding tile: zeros out (interleaved), nothing to write (planar)
    if (!C_PLANAR) {
      const int rows = min(64, t.M - m0), cols = min(NT, t.N - n0);
      for (int idx = tid; idx < rows * cols * 4; idx += MX_PROD + 64) {
        const int q = idx & 3, n = (idx >> 2) % cols, m = (idx >> 2) / cols;
        reinterpret_cast<float4*>(t.c2 + (int64_t)(m0 + m) * t.c_sm + (int64_t)(n0 + n) * t.c_sn + p0)[q] =
            make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], MX_PROD / 32);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < R; ++i) {
      mbar_init(&rawfull[i], 1);
      mbar_init(&rawempty[i], MX_PROD / 32);
    }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == MX_PROD / 32 + 1 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b0);
    if (B_PLANAR) tma_prefetch_desc(&tm_b1);
  }
  if (warp == MX_PROD / 32) tmem_alloc<TCOLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int nch = (t.K + MX_KC - 1) / MX_KC;
  const uint32_t smem0 = smem_u32(mx_smem);

  if (warp == MX_PROD / 32 + 1) {
    // ===================== TMA producer: raw chunks, the mode axis innermost (32 / 64-byte box rows) =====================
    if (lane == 0) {
      for (int c = 0; c < nch; ++c) {
        const int rs = c % R, k0 = c * MX_KC;
        if (c >= R) mbar_wait_backoff(&rawempty[rs], (uint32_t)((c / R - 1) & 1));
        uint8_t* dst = mx_smem + RAW_OFF + (size_t)rs * RAW;
        mbar_expect_tx(&rawfull[rs], RAW);
        if (t.a_k_inner) tma_load_3d(dst, &tm_a, &rawfull[rs], (int)(2 * p0), k0, m0);
        else tma_load_3d(dst, &tm_a, &rawfull[rs], (int)(2 * p0), m0, k0);
        if (B_PLANAR) {
          const int c1 = t.b_k_inner ? k0 : n0, c2 = t.b_k_inner ? n0 : k0;
          tma_load_4d(dst + MX_RAW_A, &tm_b0, &rawfull[rs], (int)loc0, c1, c2, g0);
          tma_load_4d(dst + MX_RAW_A + NT * MX_KC * 32, &tm_b1, &rawfull[rs], (int)loc0, c1, c2, g0);
        } else {
          tma_load_3d(dst + MX_RAW_A, &tm_b0, &rawfull[rs], (int)(2 * p0), n0, k0);
        }
      }
    }
  } else if (warp < MX_PROD / 32) {
    // ===================== transposers: raw chunk -> UMMA K-major tiles (tf32-rounded, re / im stacked) =====================
    for (int c = 0; c < nch; ++c) {
      const int rs = c % R, s = c % NS;
      mbar_wait(&rawfull[rs], (uint32_t)((c / R) & 1));
      if (c >= NS) mbar_wait(&empty[s], (uint32_t)((c / NS - 1) & 1));
      const uint32_t ra = smem0 + RAW_OFF + (uint32_t)rs * RAW, rb = ra + MX_RAW_A;
      const uint32_t sa = smem0 + (uint32_t)s * STAGE, sb = sa + B_OFF;
#pragma unroll
      for (int it = 0; it < 64 * MX_KC * 4 / MX_PROD; ++it) {
        const int idx = it * MX_PROD + tid;
        const int j = idx & 3, k = (idx >> 2) & 3, m = mx_row(idx >> 4);
        if (m0 + m >= t.M) continue;  // D row r depends on A' row r only, and rows past M are never read back
        float v[4];
        ld_shared_v4(ra + (uint32_t)m * t.a_rm + (uint32_t)k * t.a_rk + (uint32_t)j * 16, v);
        if (j >= 2 && !gv1) v[0] = v[1] = v[2] = v[3] = 0.f;  // padding modes: exact zeros whatever the other operand holds
        // rows 2m (re) and 2m + 1 (im) of the A' tiles of modes 2j and 2j + 1
        const uint32_t o = sa + (uint32_t)(2 * j) * MX_TILE_A + (uint32_t)(m >> 2) * MX_SBO + (uint32_t)((2 * m) & 7) * 16 + (uint32_t)k * 4;
        const float re0 = round_tf32(v[0]), im0 = round_tf32(v[1]);
        const float re1 = round_tf32(v[2]), im1 = round_tf32(v[3]);
        st_shared_f32(o, re0);
        st_shared_f32(o + MX_LBO, CONJ_B ? im0 : -im0);
        st_shared_f32(o + 16, im0);
        st_shared_f32(o + 16 + MX_LBO, CONJ_B ? -re0 : re0);
        st_shared_f32(o + MX_TILE_A, re1);
        st_shared_f32(o + MX_TILE_A + MX_LBO, CONJ_B ? im1 : -im1);
        st_shared_f32(o + MX_TILE_A + 16, im1);
        st_shared_f32(o + MX_TILE_A + 16 + MX_LBO, CONJ_B ? -re1 : re1);
      }
      if (B_PLANAR) {
#pragma unroll
        for (int it = 0; it < NBP; ++it) {
          const int idx = it * MX_PROD + tid;
          if (NT * MX_KC * 2 < MX_PROD && idx >= NT * MX_KC * 2) continue;  // (NT = 32: half the threads have no piece)
          // a quarter warp reads 128 contiguous bytes of the raw chunk whichever of (n, k) is its faster dimension; the
          // transposing stores hit 32 distinct banks under either assignment of the lane bits
          const int h = idx & 1;
          const int k = t.b_k_inner ? (idx >> 1) & 3 : (idx >> 3) & 3;
          const int n = t.b_k_inner ? idx >> 3 : ((idx >> 1) & 3) | ((idx >> 5) << 2);
          if (n0 + n >= t.N) continue;  // (likewise D column n and B' row n)
          float vr[4], vi[4];
          const uint32_t r = rb + (uint32_t)n * t.b_rn + (uint32_t)k * t.b_rk + (uint32_t)h * 16;
          ld_shared_v4(r, vr);
          ld_shared_v4(r + NT * MX_KC * 32, vi);
          const uint32_t o = sb + (uint32_t)(4 * h) * TILE_B + (uint32_t)(n >> 3) * MX_SBO + (uint32_t)(n & 7) * 16 + (uint32_t)k * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            st_shared_f32(o + (uint32_t)q * TILE_B, round_tf32(vr[q]));
            st_shared_f32(o + (uint32_t)q * TILE_B + MX_LBO, round_tf32(vi[q]));
          }
        }
      } else {
#pragma unroll
        for (int it = 0; it < NBI; ++it) {
          const int idx = it * MX_PROD + tid;
          const int j = idx & 3, k = (idx >> 2) & 3, n = idx >> 4;
          if (n0 + n >= t.N) continue;
          float v[4];
          ld_shared_v4(rb + (uint32_t)n * t.b_rn + (uint32_t)k * t.b_rk + (uint32_t)j * 16, v);
          const uint32_t o = sb + (uint32_t)(2 * j) * TILE_B + (uint32_t)(n >> 3) * MX_SBO + (uint32_t)(n & 7) * 16 + (uint32_t)k * 4;
          st_shared_f32(o, round_tf32(v[0]));
          st_shared_f32(o + MX_LBO, round_tf32(v[1]));
          st_shared_f32(o + TILE_B, round_tf32(v[2]));
          st_shared_f32(o + TILE_B + MX_LBO, round_tf32(v[3]));
        }
      }
      fence_proxy_async();  // the tensor core reads what this thread wrote through the async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full[s]);      // one arrive per warp
        mbar_arrive(&rawempty[rs]);
      }
    }

    // ===================== epilogue: TMEM -> registers -> 32-byte runs along the mode axis =====================
    mbar_wait(&done, 0);
    tc_fence_after();
    const int quad = warp & 3, half = warp >> 2;  // TMEM lane quarter, column quarter
    const int m = m0 + 16 * quad + (lane >> 1);
    const int part = lane & 1;  // 0: this lane holds Re C(m, :), 1: Im C(m, :)
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int nb = half * (NT / 4); nb < (half + 1) * (NT / 4); nb += 8) {
      float v[MX_PT][8];
#pragma unroll
      for (int q = 0; q < MX_PT; ++q) tmem_ld8(trow + (uint32_t)(q * NT + nb), v[q]);
      tmem_ld_wait();
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) {
        const int n = n0 + nb + nn;
        const bool ok = m < t.M && n < t.N;
        if (C_PLANAR) {
          // per store instruction a lane pair fills one whole 32-byte sector of ONE plane: the re lane writes modes 0..3 of both
          // planes, the im lane modes 4..7; each sends the partner the four values it needs
          float o[4], r[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[q] = part ? v[4 + q][nn] : v[q][nn];
            r[q] = __shfl_xor_sync(0xffffffffu, part ? v[q][nn] : v[4 + q][nn], 1);
          }
          const int64_t off = (int64_t)m * t.c_sm + (int64_t)n * t.c_sn + woff + 4 * part;
          if (ok && (part == 0 || gv1)) {
            *reinterpret_cast<float4*>(t.cr + off) = part ? make_float4(r[0], r[1], r[2], r[3]) : make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(t.ci + off) = part ? make_float4(o[0], o[1], o[2], o[3]) : make_float4(r[0], r[1], r[2], r[3]);
          }
        } else {
          // per store instruction a lane pair fills one whole 32-byte sector: the re lane writes modes 0,1 / 4,5, the im lane
          // modes 2,3 / 6,7; each sends the partner the four values it needs
          float x[4], y[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int qm = (q & 1) + 4 * (q >> 1);  // 0, 1, 4, 5: the re lane's modes; the im lane's are qm + 2
            const float send = part ? v[qm][nn] : v[qm + 2][nn];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            x[q] = round_tf32(part ? recv : v[qm][nn]);
            y[q] = round_tf32(part ? v[qm + 2][nn] : recv);
          }
          if (ok) {
            float4* dst = reinterpret_cast<float4*>(t.c2 + (int64_t)m * t.c_sm + (int64_t)n * t.c_sn + p0 + 2 * part);
            dst[0] = make_float4(x[0], y[0], x[1], y[1]);
            dst[2] = make_float4(x[2], y[2], x[3], y[3]);
          }
        }
      }
    }
  } else if (warp == MX_PROD / 32) {
    // ===================== MMA issuer (warp convergent, one elected lane issues) =====================
    const uint32_t idesc = idesc_tf32(128, NT, /*a_mn_major=*/false, /*b_mn_major=*/false);
    const uint64_t adesc0 = smem_desc_none(smem0, MX_LBO, MX_SBO);
    const uint64_t bdesc0 = smem_desc_none(smem0 + B_OFF, MX_LBO, MX_SBO);
    const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
    for (int c = 0; c < nch; ++c) {
      const int s = c % NS;
      mbar_wait(&full[s], (uint32_t)((c / NS) & 1));
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < MX_PT; ++q)
          mma_tf32_lohi(tmem_base + (uint32_t)(q * NT), a_lo0 + (((uint32_t)s * STAGE + (uint32_t)q * MX_TILE_A) >> 4), a_hi,
                        b_lo0 + (((uint32_t)s * STAGE + (uint32_t)q * TILE_B) >> 4), b_hi, idesc, c ? 1u : 0u);
        mma_commit(&empty[s]);
        if (c == nch - 1) mma_commit(&done);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MX_PROD / 32) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

struct MixMaps {
  CUtensorMap a, b0, b1;
};

template <int NT, bool B_PLANAR, bool C_PLANAR, bool CONJ_B>
static int launch_nt(MixArgs t, int64_t J, cudaStream_t s) {
  constexpr int NS = 2;
  constexpr int R = NT == 32 ? 5 : 3;
  constexpr size_t smem = (size_t)NS * mx_stage_bytes<NT>() + (size_t)R * mx_raw_bytes<NT>();
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static DeviceOnce attr;
  if (attr.first_use())
    cudaFuncSetAttribute(mode_gemm_tc_kernel<NT, NS, R, B_PLANAR, C_PLANAR, CONJ_B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
  // tensor maps: the mode axis is the contiguous one of every operand; (row, k) ordered by their strides
  MixMaps mp;
  {
    t.a_k_inner = t.a_sk < t.a_sm ? 1 : 0;
    const uint64_t dims[3] = {(uint64_t)(2 * J), (uint64_t)(t.a_k_inner ? t.K : t.M), (uint64_t)(t.a_k_inner ? t.M : t.K)};
    const uint64_t str[2] = {(uint64_t)(t.a_k_inner ? t.a_sk : t.a_sm) * 8, (uint64_t)(t.a_k_inner ? t.a_sm : t.a_sk) * 8};
    const uint32_t box[3] = {16, (uint32_t)(t.a_k_inner ? MX_KC : 64), (uint32_t)(t.a_k_inner ? 64 : MX_KC)};
    PDEQX_TRY(make_tensor_map(&mp.a, t.a, 3, dims, str, box, TM_NONE));
    t.a_rk = t.a_k_inner ? 64 : 64 * 64;
    t.a_rm = t.a_k_inner ? 64 * MX_KC : 64;
  }
  if (B_PLANAR) {
    // weights (G, o, i, *m): n = o, k = i (mix) or n = i, k = o (adjoint mix)
    t.b_k_inner = t.b_sk < t.b_sn ? 1 : 0;
    const int64_t mprod = t.mm.mprod;
    const int64_t d1 = (t.b_k_inner ? t.b_sn : t.b_sk) / mprod;  // = I
    const int64_t d2 = t.w_corner / (d1 * mprod);                // = O
    int G = 1;
    for (int a = 0; a < t.mm.D - 1; ++a) G *= 2;
    const uint64_t dims[4] = {(uint64_t)mprod, (uint64_t)d1, (uint64_t)d2, (uint64_t)G};
    const uint64_t str[3] = {(uint64_t)mprod * 4, (uint64_t)d1 * mprod * 4, (uint64_t)t.w_corner * 4};
    const uint32_t box[4] = {8, (uint32_t)(t.b_k_inner ? MX_KC : NT), (uint32_t)(t.b_k_inner ? NT : MX_KC), 1};
    PDEQX_TRY(make_tensor_map(&mp.b0, t.br, 4, dims, str, box, TM_NONE));
    PDEQX_TRY(make_tensor_map(&mp.b1, t.bi, 4, dims, str, box, TM_NONE));
    t.b_rk = t.b_k_inner ? 32 : NT * 32;
    t.b_rn = t.b_k_inner ? 32 * MX_KC : 32;
  } else {
    // interleaved B (dW): x_hat [b = k, i = n, J]
    t.b_k_inner = 0;
    const uint64_t dims[3] = {(uint64_t)(2 * J), (uint64_t)t.N, (uint64_t)t.K};
    const uint64_t str[2] = {(uint64_t)t.b_sn * 8, (uint64_t)t.b_sk * 8};
    const uint32_t box[3] = {16, (uint32_t)NT, (uint32_t)MX_KC};
    PDEQX_TRY(make_tensor_map(&mp.b0, t.b2, 3, dims, str, box, TM_NONE));
    mp.b1 = mp.b0;
    t.b_rk = NT * 64;
    t.b_rn = 64;
  }
  dim3 grid((unsigned)(J / MX_PT), (unsigned)ceil_div(t.N, NT), (unsigned)ceil_div(t.M, 64));
  PDEQX_CUDA(launch_pdl(mode_gemm_tc_kernel<NT, NS, R, B_PLANAR, C_PLANAR, CONJ_B>, grid, dim3(MX_PROD + 64), smem, s, mp.a, mp.b0, mp.b1, t));
  return PDEQX_OK;
}

template <bool B_PLANAR, bool C_PLANAR, bool CONJ_B>
static int launch(const MixArgs& t, int64_t J, cudaStream_t s) {
  // 32-column tiles unless 64-column tiles already fill the GPU
  const int64_t tiles64 = (J / MX_PT) * ceil_div(t.N, 64) * ceil_div(t.M, 64);
  if (t.N <= 32 || tiles64 < 148) return launch_nt<32, B_PLANAR, C_PLANAR, CONJ_B>(t, J, s);
  return launch_nt<64, B_PLANAR, C_PLANAR, CONJ_B>(t, J, s);
}

}  // namespace tc

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

bool tc_mode_gemm_available(const pdeqx_spectral_plan* p) {
  // experiment knob: PDEQX_TC_MODEMIX=0 keeps the CUDA-core kernel, =2 takes this one whenever the shape allows it
  static const int knob = getenv("PDEQX_TC_MODEMIX") ? atoi(getenv("PDEQX_TC_MODEMIX")) : 1;
  if (knob == 0 || p->desc.precision != PDEQX_PREC_TF32 || !g_fast_paths.load(std::memory_order_relaxed) || !p->tc) return false;
  // measured (B200, profiles/r2_modemix.md): ahead of the CUDA-core kernel from 16 samples per GPU and 64 channels on; level with
  // it where that kernel is a pure stream over the weights (8 samples per GPU), behind it on the padded 32-channel mode space
  // of configs[4]
  if (knob != 2 && (p->desc.batch < 16 || p->desc.c_in < 64 || p->desc.c_out < 64)) return false;
  const int D = p->desc.ndim;
  int64_t mprod = 1;
  for (int a = 0; a < D; ++a) mprod *= p->desc.modes[a];
  // 8-mode tiles must not straddle a row of the data layout; 16-byte weight runs start on 4-mode boundaries
  return p->mlast % 8 == 0 && p->desc.modes[D - 1] % 4 == 0 && mprod % 4 == 0 && p->J % 8 == 0 && p->J / 8 < (1ll << 31) &&
         ceil_div(p->desc.batch, 64) <= 65535;
}

int tc_mode_mix(const pdeqx_spectral_plan* p, const float* in, const float* w_re, const float* w_im, bool transposed_conj,
                float* out, cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  const int O = d.c_out, I = d.c_in;
  const int U = transposed_conj ? I : O, V = transposed_conj ? O : I;
  PDEQX_CHECK_ARG(aligned16(in) && aligned16(out) && aligned16(w_re) && aligned16(w_im), "tc_mode_mix: pointers must be 16-byte aligned");
  tc::MixArgs t{};
  t.mm = make_map(p);
  t.a = (const float2*)in; t.a_sm = (int64_t)V * p->J; t.a_sk = p->J;
  t.br = w_re; t.bi = w_im;
  // W(u,v): plain -> w[o=u, i=v]; transposed_conj -> conj(w[o=v, i=u])
  t.b_sn = transposed_conj ? t.mm.mprod : (int64_t)I * t.mm.mprod;
  t.b_sk = transposed_conj ? (int64_t)I * t.mm.mprod : t.mm.mprod;
  t.c2 = (float2*)out; t.c_sm = (int64_t)U * p->J; t.c_sn = p->J;
  t.M = d.batch; t.N = U; t.K = V; t.w_corner = (int64_t)O * I * t.mm.mprod;
  if (transposed_conj) PDEQX_TRY((tc::launch<true, false, true>(t, p->J, s)));
  else PDEQX_TRY((tc::launch<true, false, false>(t, p->J, s)));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

int tc_mode_mix_dw(const pdeqx_spectral_plan* p, const float* ghat, const float* xhat, float* gw_re, float* gw_im,
                   cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  PDEQX_CHECK_ARG(aligned16(ghat) && aligned16(xhat) && aligned16(gw_re) && aligned16(gw_im), "tc_mode_mix_dw: pointers must be 16-byte aligned");
  tc::MixArgs t{};
  t.mm = make_map(p);
  // M = o, N = i, K = b: gw[o,i] = sum_b ghat[b,o] conj(xhat[b,i])
  t.a = (const float2*)ghat; t.a_sm = p->J; t.a_sk = (int64_t)d.c_out * p->J;
  t.b2 = (const float2*)xhat; t.b_sn = p->J; t.b_sk = (int64_t)d.c_in * p->J;
  t.cr = gw_re; t.ci = gw_im; t.c_sm = (int64_t)d.c_in * t.mm.mprod; t.c_sn = t.mm.mprod;
  t.M = d.c_out; t.N = d.c_in; t.K = d.batch; t.w_corner = (int64_t)d.c_out * d.c_in * t.mm.mprod;
  PDEQX_TRY((tc::launch<false, true, true>(t, p->J, s)));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx
