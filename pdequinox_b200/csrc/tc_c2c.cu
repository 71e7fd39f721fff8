// Leading-axis (complex-to-complex) truncated DFT stages of SpectralConv on tcgen05 (TF32).
//
//   out[outer, r, :] = sum_h T[r, h] * in[outer, h, :]        T complex [n_out, n_in], rows ':' complex-interleaved
//
// These are the stages between the first-stage r2c DFT and the per-mode channel mix (forward: s_a -> 2 m_a rows) and
// between the mix and the last-stage slab kernel (inverse: 2 m_a -> s_a rows); reference: the leading axes of
// jnp.fft.rfftn / irfftn in pdequinox/conv/_spectral_conv.py:131,136, restricted to the retained modes.
//
// GEMM form (no complex arithmetic on the tensor core): with the interleaved (re, im) floats of a row as the M index,
//     P[m, r] = sum_h in[h, m] Re T[r, h],   Q[m, r] = sum_h in[h, m] Im T[r, h]          (one GEMM, N = 2 n_out)
//     out[r, (k, re)] = P[(k, re), r] - Q[(k, im), r],   out[r, (k, im)] = P[(k, im), r] + Q[(k, re), r]
// so the epilogue only needs the partner lane (lane ^ 1). M = 128 = four 32-float row chunks ("blocks") of possibly
// different outer indices; A = in (MN-major: the 32 floats of a row chunk are contiguous -> TMA SWIZZLE_128B_ATOM_32B),
// B = [Re T ; Im T] (K-major, resident in shared memory), K = n_in streamed in chunks of up to 64 rows.
#include <string.h>

#include <vector>

#include "spectral_plan.h"
#include "tc_common.cuh"

namespace pdeqx {
namespace tc {

struct C2cParams {
  int n_in, n_out, N;       // N = 2 n_part accumulator columns per unit
  int n_part, parts;        // output rows per unit; units per block group (n_out = n_part * parts)
  int kchunk, kchunks;      // rows per TMA stage, stages per unit
  int q_per_outer;          // 32-float chunks per row
  int stages, acc_stages;
  int round_out;            // round the outputs to tf32 (nearest): the consumer is another tcgen05 stage, which would truncate
  int64_t blocks, units;    // M-blocks (outer x chunk), units of 4 blocks
  int64_t inner_f;          // floats per row
  float* out;
};

constexpr int kC2cThreads = 384;  // warps 0-2 roles, 4-11 epilogue (two warps per TMEM lane quarter)

__global__ void __launch_bounds__(kC2cThreads, 1)
c2c_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_tab, C2cParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t tab_chunk_bytes = (uint32_t)(2 * p.n_out) * 128;     // [2 n_out rows][32 k]
  const uint32_t tab_bytes = (uint32_t)(p.n_in / 32) * tab_chunk_bytes;
  const uint32_t box_bytes = (uint32_t)p.kchunk * 128;                 // [kchunk rows][32 floats]
  const uint32_t stage_bytes = 4 * box_bytes;
  uint8_t* stage0 = smem + tab_bytes;
  uint64_t* bars = (uint64_t*)(stage0 + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;        // [stages <= 8]
  uint64_t* empty = bars + 8;   // [stages]
  uint64_t* tfull = bars + 16;  // [2]
  uint64_t* tempty = bars + 18; // [2]
  uint64_t* tabfull = bars + 20;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 21);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages, AS = p.acc_stages;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_tab);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(tabfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int64_t first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(tabfull, tab_bytes);
      for (int c = 0; c < p.n_in / 32; ++c)
        for (int nb = 0; nb < 2 * p.n_out; nb += 256)
          tma_load_2d(smem + (size_t)c * tab_chunk_bytes + (size_t)nb * 128, &tm_tab, tabfull, c * 32, nb);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t u = first; u < p.units; u += stride) {
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait_backoff(&empty[s], ph ^ 1);
          uint8_t* dst = stage0 + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], stage_bytes);
#pragma unroll
          for (int mb = 0; mb < 4; ++mb) {
            const int64_t id = (u / p.parts) * 4 + mb;  // blocks past the end are out of bounds: zero-filled, never stored
            const int outer = (int)(id / p.q_per_outer), q = (int)(id % p.q_per_outer);
            tma_load_3d(dst + mb * box_bytes, &tm_in, &full[s], q * 32, kc * p.kchunk, outer);
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp convergent, one elected lane issues) =====================
    // one MMA with N = 2 n_out when a unit covers every output row ([Re T ; Im T] rows are contiguous), else two MMAs
    // (Re rows of the part, Im rows of the part) with N = n_part each
    const bool single = p.parts == 1;
    const uint32_t idesc = idesc_tf32(128, single ? p.N : p.n_part, /*a_mn_major=*/true, /*b_mn_major=*/false);
    const uint64_t adesc0 = smem_desc_sw128_atom32(smem_u32(stage0), box_bytes, 512);
    const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem), 0, 1024);
    const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
    const uint32_t stage16 = stage_bytes >> 4, tchunk16 = tab_chunk_bytes >> 4;
    const int ksteps = p.kchunk / 8;
    mbar_wait(tabfull, 0);
    int s = 0, a = 0;
    uint32_t ph = 0, aph = 0;
    for (int64_t u = first; u < p.units; u += stride) {
      const int part = (int)(u % p.parts);
      const uint32_t re16 = (uint32_t)(part * p.n_part) * 8, im16 = (uint32_t)(p.n_out + part * p.n_part) * 8;  // rows * 128 B / 16
      mbar_wait(&tempty[a], aph ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(a * p.N);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)s * stage16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kg = kc * ksteps + ks;  // global k-step: table chunk kg >> 2, 32-byte slice kg & 3
            const uint32_t b_lo = b_lo0 + (uint32_t)(kg >> 2) * tchunk16 + (uint32_t)(kg & 3) * 2;
            const uint32_t acc = (kc | ks) ? 1u : 0u;
            mma_tf32_lohi(d0, a_lo + (uint32_t)ks * 64, a_hi, b_lo + re16, b_hi, idesc, acc);
            if (!single) mma_tf32_lohi(d0 + (uint32_t)p.n_part, a_lo + (uint32_t)ks * 64, a_hi, b_lo + im16, b_hi, idesc, acc);
          }
          mma_commit(&empty[s]);
          if (kc == p.kchunks - 1) mma_commit(&tfull[a]);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (++a == AS) { a = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: combine re/im parts, store rows =====================
    const int q4 = warp & 3;            // which of the unit's four blocks (= TMEM lane quarter)
    const int rh = (warp - 4) >> 2;     // which half of the unit's output rows
    const int rows_w = p.n_part / 2;    // rows per warp
    int a = 0;
    uint32_t aph = 0;
    for (int64_t u = first; u < p.units; u += stride) {
      const int part = (int)(u % p.parts);
      const int64_t id = (u / p.parts) * 4 + q4;
      const bool live = id < p.blocks;
      const int64_t outer = id / p.q_per_outer;
      const int q = (int)(id % p.q_per_outer);
      const int row0 = part * p.n_part + rh * rows_w;
      // (warp-uniform) output pointer of this lane's column; consecutive output rows are inner_f floats apart -- the
      // per-row address is a pointer increment, not a 64-bit multiply
      float* rowp = p.out + (outer * (int64_t)p.n_out + row0) * p.inner_f + (int64_t)q * 32 + lane;
      const int64_t rstride = p.inner_f;
      const bool odd = lane & 1;
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * p.N + rh * rows_w);
      for (int r0 = 0; r0 < rows_w; r0 += 16) {
        float pr[16], qi[16];
        tmem_ld16(t0 + (uint32_t)r0, pr);
        tmem_ld16(t0 + (uint32_t)(p.n_part + r0), qi);
        tmem_ld_wait();
        if (r0 + 16 >= rows_w) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[a]);
        }
        if (live) {  // warp-uniform: the whole 32-float block is inside the tensor or not
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float other = __shfl_xor_sync(0xffffffffu, qi[j], 1);
            float v = odd ? pr[j] + other : pr[j] - other;
            if (p.round_out) v = round_tf32(v);
            *rowp = v;
            rowp += rstride;
          }
        }
      }
      if (++a == AS) { a = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// plan-side tables: for every leading axis and each of the four table kinds, [Re T ; Im T] as [2 n_out][n_in]
// ------------------------------------------------------------------------------------------
struct C2cTable {
  float* dev = nullptr;
  CUtensorMap map;
  int n_in = 0, n_out = 0;
  bool ok = false;
};
struct TcC2cPlan {
  C2cTable t[PDEQX_MAX_DIMS][4];  // [axis][kind]: 0 fwd, 1 inv, 2 inv_adj, 3 fwd_adj
  int sms = 148;
};

static float host_round_tf32_c2c(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return v;
  u += 0x1000u;
  u &= 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

int tc_c2c_plan_init(pdeqx_spectral_plan* p) {
  p->tc_c2c = nullptr;
  const pdeqx_spectral_desc& d = p->desc;
  if (d.precision != PDEQX_PREC_TF32 || d.ndim < 2) return PDEQX_OK;
  int dev = 0, major = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return PDEQX_OK;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10 || !tc::encode_tiled_fn()) return PDEQX_OK;
  TcC2cPlan* t = new TcC2cPlan();
  t->sms = sms;
  for (int a = 0; a < d.ndim - 1; ++a) {
    const int s = d.spatial[a], m2 = p->n2[a];
    const float2* src[4] = {p->t_fwd[a], p->t_inv[a], p->t_inv_adj[a], p->t_fwd_adj[a]};
    const int n_out[4] = {m2, s, m2, s}, n_in[4] = {s, m2, s, m2};
    for (int k = 0; k < 4; ++k) {
      C2cTable& tb = t->t[a][k];
      tb.n_in = n_in[k];
      tb.n_out = n_out[k];
      // shape limits of the kernel: K chunks of 32, N = 2 n_out <= 512 accumulator columns, 16-column epilogue steps
      if (n_in[k] % 32 != 0 || n_out[k] % 32 != 0 || (2 * n_out[k] > 256 && (2 * n_out[k]) % 256 != 0)) continue;
      if ((size_t)2 * n_out[k] * n_in[k] * 4 > 96 * 1024) continue;  // table must stay resident in shared memory
      std::vector<float2> h((size_t)n_out[k] * n_in[k]);
      if (cudaMemcpy(h.data(), src[k], h.size() * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) continue;
      std::vector<float> tt((size_t)2 * n_out[k] * n_in[k]);
      for (int r = 0; r < n_out[k]; ++r)
        for (int c = 0; c < n_in[k]; ++c) {
          tt[(size_t)r * n_in[k] + c] = host_round_tf32_c2c(h[(size_t)r * n_in[k] + c].x);
          tt[(size_t)(n_out[k] + r) * n_in[k] + c] = host_round_tf32_c2c(h[(size_t)r * n_in[k] + c].y);
        }
      if (cudaMalloc((void**)&tb.dev, tt.size() * 4) != cudaSuccess) continue;
      if (cudaMemcpy(tb.dev, tt.data(), tt.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) continue;
      uint64_t dims[2] = {(uint64_t)n_in[k], (uint64_t)2 * n_out[k]};
      uint64_t str[1] = {(uint64_t)n_in[k] * 4};
      uint32_t box[2] = {32, (uint32_t)(2 * n_out[k] < 256 ? 2 * n_out[k] : 256)};
      if (tc::make_tensor_map(&tb.map, tb.dev, 2, dims, str, box, tc::TM_SW128) != PDEQX_OK) continue;
      tb.ok = true;
    }
  }
  p->tc_c2c = t;
  return PDEQX_OK;
}

void tc_c2c_plan_free(pdeqx_spectral_plan* p) {
  TcC2cPlan* t = (TcC2cPlan*)p->tc_c2c;
  if (!t) return;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a)
    for (int k = 0; k < 4; ++k) cudaFree(t->t[a][k].dev);
  delete t;
  p->tc_c2c = nullptr;
}

bool tc_c2c_available(const pdeqx_spectral_plan* p, int axis, int kind, int64_t inner_complex) {
  TcC2cPlan* t = (TcC2cPlan*)p->tc_c2c;
  if (!t || !g_fast_paths.load(std::memory_order_relaxed)) return false;
  return t->t[axis][kind].ok && (2 * inner_complex) % 32 == 0;
}

int tc_c2c(const pdeqx_spectral_plan* p, int axis, int kind, const float* in, float* out, int64_t outer, int64_t inner_complex,
           cudaStream_t s, bool round_out) {
  TcC2cPlan* t = (TcC2cPlan*)p->tc_c2c;
  const C2cTable& tb = t->t[axis][kind];
  PDEQX_CHECK_ARG(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "tc_c2c: pointers must be 16-byte aligned");
  PDEQX_CHECK_ARG(outer < (1ll << 31), "tc_c2c: outer too large");
  tc::C2cParams cp{};
  cp.n_in = tb.n_in;
  cp.n_out = tb.n_out;
  cp.parts = (2 * tb.n_out > 256) ? (2 * tb.n_out) / 256 : 1;
  cp.n_part = tb.n_out / cp.parts;
  cp.N = 2 * cp.n_part;
  cp.kchunk = tb.n_in >= 64 && tb.n_in % 64 == 0 ? 64 : 32;
  cp.kchunks = tb.n_in / cp.kchunk;
  cp.inner_f = 2 * inner_complex;
  cp.q_per_outer = (int)(cp.inner_f / 32);
  cp.blocks = outer * cp.q_per_outer;
  cp.units = ceil_div(cp.blocks, 4) * cp.parts;
  cp.acc_stages = 2;
  cp.round_out = round_out ? 1 : 0;
  cp.out = out;
  const uint32_t tab_bytes = (uint32_t)(2 * tb.n_out) * tb.n_in * 4;
  const uint32_t stage_bytes = 4 * (uint32_t)cp.kchunk * 128;
  int stages = (int)((227 * 1024 - 2048 - tab_bytes) / stage_bytes);
  if (stages > 6) stages = 6;
  PDEQX_CHECK_ARG(stages >= 2, "tc_c2c: table leaves no room for the pipeline");
  cp.stages = stages;
  CUtensorMap tm_in;
  {
    uint64_t dims[3] = {(uint64_t)cp.inner_f, (uint64_t)tb.n_in, (uint64_t)outer};
    uint64_t str[2] = {(uint64_t)cp.inner_f * 4, (uint64_t)tb.n_in * cp.inner_f * 4};
    uint32_t box[3] = {32, (uint32_t)cp.kchunk, 1};
    PDEQX_TRY(tc::make_tensor_map(&tm_in, in, 3, dims, str, box, tc::TM_SW128_ATOM32));
  }
  const size_t smem = tab_bytes + (size_t)stages * stage_bytes + 1024 + 1024;
  static DeviceOnce attr;      
  if (attr.first_use()) {
    PDEQX_CUDA(cudaFuncSetAttribute(tc::c2c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                
  }
  const int grid = (int)(cp.units < t->sms ? cp.units : t->sms);
  PDEQX_CUDA(launch_pdl(tc::c2c_kernel, dim3(grid), dim3(tc::kC2cThreads), smem, s, tm_in, tb.map, cp));
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx
