// SpectralConv / ClassicSpectralBlock as a truncated-DFT pipeline.
//
//   replaces  pdequinox/conv/_spectral_conv.py:107-136   (rfftn -> corner einsum -> irfftn)
//             pdequinox/blocks/_classic_spectral_block.py:72-75
//
// Only the retained modes are ever computed: the forward transform is a chain of
// GEMMs against twiddle tables (real->complex on the contiguous axis first, which
// shrinks the data by s_D/(2 m_D), then complex->complex on the remaining axes),
// followed by the per-mode channel mix and the mirrored inverse chain. The zero
// filled full spectrum of the reference is never materialised.
//
// This file owns the plan (tables) and the stage sequencing; it dispatches each
// stage to the tcgen05 kernels (tc_spectral.cu) when the shape allows and to the
// fp32 CUDA-core GEMMs (gemm_generic.cu) otherwise.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "spectral_plan.h"

namespace pdeqx {

static const double kTwoPi = 6.283185307179586476925286766559;

static int64_t kappa(int j, int s, int m) { return j < m ? j : (int64_t)s - 2 * m + j; }

// cos/sin of 2 pi (k n mod s) / s in float64
static void cs(int64_t k, int64_t n, int64_t s, double* c, double* sn) {
  int64_t r = (k * n) % s;
  double ang = kTwoPi * (double)r / (double)s;
  *c = cos(ang);
  *sn = sin(ang);
}

// Where the first stage of the adjoint-of-inverse transform of the reverse pass runs (PDEQX_FUSE_R2C):
//   1 (default) = inside the bypass weight-gradient kernel, which streams the same cotangent tile (-131 us per block at
//   configs[3]); 0 = its own kernel. A third variant (inside the epilogue of the slab kernel that produces the cotangent,
//   as per-w-tile partial sums) was measured slower (+159 us in the slab kernel for -190 us) and removed.
static int fuse_r2c_mode() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("PDEQX_FUSE_R2C");
    env = e ? atoi(e) : 1;
  }
  return env;
}

template <class T>
static int upload(const std::vector<T>& h, T** d) {
  PDEQX_CUDA(cudaMalloc((void**)d, h.size() * sizeof(T)));
  PDEQX_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return PDEQX_OK;
}

}  // namespace pdeqx

using namespace pdeqx;

extern "C" int pdeqx_spectral_plan_create(const pdeqx_spectral_desc* desc, pdeqx_spectral_plan** out) {
  PDEQX_CHECK_ARG(desc && out, "null argument");
  const int D = desc->ndim;
  PDEQX_CHECK_ARG(D >= 1 && D <= PDEQX_MAX_DIMS, "ndim must be 1..3, got %d", D);
  PDEQX_CHECK_ARG(desc->batch >= 1 && desc->c_in >= 1 && desc->c_out >= 1, "batch/channels must be >= 1");
  for (int a = 0; a < D; ++a) {
    PDEQX_CHECK_ARG(desc->spatial[a] >= 1 && desc->modes[a] >= 1, "spatial/modes must be >= 1");
    if (a == D - 1)
      PDEQX_CHECK_ARG(desc->modes[a] <= desc->spatial[a] / 2 + 1,
                      "num_modes[%d]=%d exceeds s//2+1=%d on the rfft axis", a, desc->modes[a],
                      desc->spatial[a] / 2 + 1);
    else
      PDEQX_CHECK_ARG(desc->modes[a] <= desc->spatial[a], "num_modes[%d]=%d exceeds axis length %d", a,
                      desc->modes[a], desc->spatial[a]);
  }
  pdeqx_spectral_plan* p = new pdeqx_spectral_plan();
  p->desc = *desc;
  p->G = 1 << (D - 1);
  const int sD = desc->spatial[D - 1], mD = desc->modes[D - 1];

  // ---- padded mode geometry (DATA layout only; weights keep the reference's mode counts) ----
  // The tcgen05 stages want 32 real accumulator columns on the contiguous axis (16 complex modes) and multiples of 32
  // retained rows on the leading axes. Smaller mode counts (configs[4]: 12) are padded with zero table rows / columns:
  // the padded modes carry exact zeros through every stage and the channel mix skips them.
  const bool tf32 = desc->precision == PDEQX_PREC_TF32;
  const int mP = (tf32 && mD < 16 && mD % 4 == 0 && sD % 32 == 0) ? 16 : mD;
  p->mlast = mP;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    p->n2[a] = a < D - 1 ? 2 * desc->modes[a] : 1;
    if (a < D - 1 && tf32 && mP == 16 && p->n2[a] % 32 != 0 && desc->spatial[a] % 32 == 0) p->n2[a] = (p->n2[a] + 31) / 32 * 32;
  }

  // ---- last axis (real <-> complex) ----
  std::vector<float> r2c_f((size_t)sD * 2 * mP, 0.f), c2r_i((size_t)2 * mP * sD, 0.f);
  std::vector<float> r2c_b((size_t)sD * 2 * mP, 0.f), c2r_b((size_t)2 * mP * sD, 0.f);
  for (int k = 0; k < mD; ++k) {
    double c = 2.0;
    if (k == 0) c = 1.0;
    if (sD % 2 == 0 && k == sD / 2) c = 1.0;  // Nyquist bin: pocketfft c2r drops its imaginary part
    for (int n = 0; n < sD; ++n) {
      double co, si;
      cs(k, n, sD, &co, &si);
      // x_hat = sum_n x e^{-i th}
      r2c_f[(size_t)n * 2 * mP + 2 * k] = (float)co;
      r2c_f[(size_t)n * 2 * mP + 2 * k + 1] = (float)(-si);
      // y = sum_k c/s (zr cos - zi sin)
      c2r_i[(size_t)(2 * k) * sD + n] = (float)(c / sD * co);
      c2r_i[(size_t)(2 * k + 1) * sD + n] = (float)(-c / sD * si);
      // adjoints are the transposes
      r2c_b[(size_t)n * 2 * mP + 2 * k] = (float)(c / sD * co);
      r2c_b[(size_t)n * 2 * mP + 2 * k + 1] = (float)(-c / sD * si);
      c2r_b[(size_t)(2 * k) * sD + n] = (float)co;
      c2r_b[(size_t)(2 * k + 1) * sD + n] = (float)(-si);
    }
  }
  int rc = PDEQX_OK;
  if ((rc = upload(r2c_f, &p->t_r2c_fwd)) || (rc = upload(c2r_i, &p->t_c2r_inv)) ||
      (rc = upload(r2c_b, &p->t_r2c_bwd)) || (rc = upload(c2r_b, &p->t_c2r_bwd))) {
    pdeqx_spectral_plan_destroy(p);
    return rc;
  }

  // ---- other axes (complex <-> complex) ----
  for (int a = 0; a < D - 1; ++a) {
    const int s = desc->spatial[a], m = desc->modes[a], np2 = p->n2[a];  // np2 >= 2m rows / columns, the extra ones zero
    const float2 zero2 = make_float2(0.f, 0.f);
    std::vector<float2> fwd((size_t)np2 * s, zero2), inv((size_t)s * np2, zero2), inv_adj((size_t)np2 * s, zero2),
        fwd_adj((size_t)s * np2, zero2);
    for (int j = 0; j < 2 * m; ++j) {
      // overlap (2m > s): the later (high) corner block overwrites the low one in the reference
      double keep = 1.0;
      if (j < m) {
        for (int jj = m; jj < 2 * m; ++jj)
          if (kappa(jj, s, m) == kappa(j, s, m)) keep = 0.0;
      }
      for (int n = 0; n < s; ++n) {
        double co, si;
        cs(kappa(j, s, m), n, s, &co, &si);
        fwd[(size_t)j * s + n] = make_float2((float)co, (float)(-si));
        fwd_adj[(size_t)n * np2 + j] = make_float2((float)co, (float)si);
        inv[(size_t)n * np2 + j] = make_float2((float)(keep * co / s), (float)(keep * si / s));
        inv_adj[(size_t)j * s + n] = make_float2((float)(keep * co / s), (float)(-keep * si / s));
      }
    }
    if ((rc = upload(fwd, &p->t_fwd[a])) || (rc = upload(inv, &p->t_inv[a])) ||
        (rc = upload(inv_adj, &p->t_inv_adj[a])) || (rc = upload(fwd_adj, &p->t_fwd_adj[a]))) {
      pdeqx_spectral_plan_destroy(p);
      return rc;
    }
  }

  // ---- sizes ----
  int64_t lead = 1;  // prod s_a, a < D-1
  int64_t J = mP;    // (padded) retained modes per channel: prod n2_a (a < D-1) * m_D
  for (int a = 0; a < D - 1; ++a) {
    lead *= desc->spatial[a];
    J *= p->n2[a];
  }
  p->lead = lead;
  p->J = J;
  const int64_t cmax = desc->c_in > desc->c_out ? desc->c_in : desc->c_out;
  // largest intermediate: right after the r2c stage / right before the c2r stage
  int64_t lead_max = 1;  // with overlapping slices (2m > s) a truncated axis can be longer than the full one
  for (int a = 0; a < D - 1; ++a) lead_max *= (desc->spatial[a] > p->n2[a] ? desc->spatial[a] : p->n2[a]);
  p->stage_floats = 2 * (int64_t)desc->batch * cmax * lead_max * mP;
  p->mode_floats = (2 * (int64_t)desc->batch * cmax * J + 63) / 64 * 64;
  // + the per-call scratch of the tcgen05 stages (two re-laid-out bypass weights, folded lifting factors): the plan itself
  // stays immutable, so one plan can serve several streams as long as each call brings its own workspace
  p->small_floats = 2 * cmax * cmax + 512;
  p->ws_bytes = (size_t)(2 * p->stage_floats + 2 * p->mode_floats + p->small_floats) * sizeof(float) + 256;
  rc = tc_plan_init(p);  // builds the tcgen05-side operand tables when the shape qualifies
  if (rc == PDEQX_OK) rc = tc_c2c_plan_init(p);
  if (rc != PDEQX_OK) {
    pdeqx_spectral_plan_destroy(p);
    return rc;
  }
  *out = p;
  return PDEQX_OK;
}

extern "C" void pdeqx_spectral_plan_destroy(pdeqx_spectral_plan* p) {
  if (!p) return;
  cudaFree(p->t_r2c_fwd);
  cudaFree(p->t_c2r_inv);
  cudaFree(p->t_r2c_bwd);
  cudaFree(p->t_c2r_bwd);
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    cudaFree(p->t_fwd[a]);
    cudaFree(p->t_inv[a]);
    cudaFree(p->t_inv_adj[a]);
    cudaFree(p->t_fwd_adj[a]);
  }
  tc_plan_free(p);
  tc_c2c_plan_free(p);
  delete p;
}

extern "C" size_t pdeqx_spectral_workspace_bytes(const pdeqx_spectral_plan* p) { return p ? p->ws_bytes : 0; }
extern "C" size_t pdeqx_spectral_xhat_floats(const pdeqx_spectral_plan* p) {
  return p ? (size_t)(2 * (int64_t)p->desc.batch * p->desc.c_in * p->J) : 0;
}
extern "C" size_t pdeqx_spectral_z_floats(const pdeqx_spectral_plan* p) {
  if (!p) return 0;
  return (size_t)(2 * (int64_t)p->desc.batch * p->desc.c_out * p->lead * p->mlast);
}

extern "C" int pdeqx_spectral_plan_tcgen05_stages(const pdeqx_spectral_plan* p) { return p ? tc_stage_mask(p) : 0; }

namespace pdeqx {

// real [R, s_D] -> complex [R, m_D]
static int stage_r2c(const pdeqx_spectral_plan* p, const float* x, const float* table, int64_t rows, float* out,
                     cudaStream_t s) {
  const int D = p->desc.ndim;
  const int sD = p->desc.spatial[D - 1], mD = p->mlast;
  TimedScope timed(PDEQX_K_SPECTRAL_R2C, s);
  if (tc_r2c_available(p)) return tc_r2c(p, x, table == p->t_r2c_bwd, rows, out, s);
  SgemmArgs a{};
  a.A = x; a.B = table; a.C = out;
  a.M = rows; a.N = 2 * mD; a.K = sD;
  a.lda_m = sD; a.lda_k = 1; a.ldb = 2 * mD; a.ldc = 2 * mD;
  a.batch = 1; a.act = PDEQX_ACT_IDENTITY;
  return sgemm_batched(a, s);
}

// complex [R, m_D] -> real [R, s_D]; out = act(result + add)
static int stage_c2r(const pdeqx_spectral_plan* p, const float* z, const float* table, int64_t rows, const float* add,
                     int act, float* out, cudaStream_t s) {
  const int D = p->desc.ndim;
  const int sD = p->desc.spatial[D - 1], mD = p->mlast;
  TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
  SgemmArgs a{};
  a.A = z; a.B = table; a.C = out; a.add = add;
  a.M = rows; a.N = sD; a.K = 2 * mD;
  a.lda_m = 2 * mD; a.lda_k = 1; a.ldb = sD; a.ldc = sD;
  a.batch = 1; a.act = act;
  timed_bytes(4.0 * (double)rows * (2.0 * mD + sD + (add ? sD : 0)));
  return sgemm_batched(a, s);
}

// One complex stage along axis `a` of a tensor [B*C, n_0.., n_a, .., inner] (complex).
//   n_in -> n_out through table [n_out x n_in].
static int stage_c2c(const pdeqx_spectral_plan* p, int axis, int kind, const float2* table, const float* in, float* out,
                     int64_t outer, int n_in, int n_out, int64_t inner, cudaStream_t s) {
  TimedScope timed(PDEQX_K_SPECTRAL_MODES, s);
  // forward-type stages (kinds 0, 2) walk the axes downwards and end in the fp32 channel mix; inverse-type stages (1, 3)
  // end in the tcgen05 slab kernel: every output that another tensor-core stage will read is rounded to tf32 here
  const bool feeds_tc = (kind == 0 || kind == 2) ? axis > 0 : true;
  if (tc_c2c_available(p, axis, kind, inner)) return tc_c2c(p, axis, kind, in, out, outer, inner, s, feeds_tc);
  return cgemm_shared_a(table, (const float2*)in, (float2*)out, outer, n_out, n_in, inner, s);
}


__device__ __forceinline__ float to_tf32(float v) {  // round to nearest, as the tcgen05 stages expect from their producers
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__global__ void fill_kernel(float* __restrict__ out, int64_t n, const float* __restrict__ value) {
  pdl_enter();
  const float v = value ? value[0] : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}

// ---- a pointwise 1 -> C lifting in front of the block, folded in (the lifted tensor x = w (x) u + b never exists) ----
// a[o] = sum_i Wb[o,i] wl[i];  c[o] = sum_i Wb[o,i] bl[i] + bb[o]: the bypass of the lifted input is a[o] u[pixel] + c[o]
__global__ void lift_fold_bypass_kernel(const float* __restrict__ wb, const float* __restrict__ bb, const float* __restrict__ wl,
                                        const float* __restrict__ bl, int co, int ci, float* __restrict__ a, float* __restrict__ c) {
  pdl_enter();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= co) return;
  float sa = 0.f, sc = bb ? bb[o] : 0.f;
  for (int i = 0; i < ci; ++i) {
    const float w = wb[(int64_t)o * ci + i];
    sa = fmaf(w, wl[i], sa);
    if (bl) sc = fmaf(w, bl[i], sc);
  }
  a[o] = sa;
  c[o] = sc;
}
// first DFT stage of the lifted input from the first DFT stage U1 of the field (the transform is linear):
//   X1[(b C + o) lead + l][k] = wl[o] U1[b lead + l][k] + [k == 0] bl[o] s_D     (sum_w cos(0) = s_D; every other column sums to 0)
__global__ void __launch_bounds__(256)
lift_x1_kernel(const float4* __restrict__ u1, const float* __restrict__ wl, const float* __restrict__ bl, float sD, int C, int64_t lead,
               int64_t total4, int round, float4* __restrict__ x1) {
  pdl_enter();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k4 = (int)(idx & 7);          // float4 index inside the 32-float row
    const int64_t row = idx >> 3;            // (b C + o) lead + l
    const int64_t l = row % lead;
    const int64_t bo = row / lead;
    const int o = (int)(bo % C);
    const int64_t b = bo / C;
    float4 v = __ldg(u1 + ((b * lead + l) << 3) + k4);
    const float w = wl[o];
    v.x *= w; v.y *= w; v.z *= w; v.w *= w;
    if (k4 == 0 && bl) v.x = fmaf(bl[o], sD, v.x);
    if (round) {
      v.x = to_tf32(v.x);
      v.y = to_tf32(v.y);
      v.z = to_tf32(v.z);
      v.w = to_tf32(v.w);
    }
    x1[idx] = v;
  }
}
// bypass gradients from the two reductions s1[o] = sum g u, s0[o] = sum g:  gW[o,i] = s1[o] wl[i] + s0[o] bl[i];  gb[o] = s0[o]
__global__ void lift_unfold_wgrad_kernel(const float* __restrict__ s1, const float* __restrict__ s0, const float* __restrict__ wl,
                                         const float* __restrict__ bl, int co, int ci, float* __restrict__ gw, float* __restrict__ gb) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= co * ci) return;
  const int o = idx / ci, i = idx % ci;
  gw[idx] = s1[o] * wl[i] + (bl ? s0[o] * bl[i] : 0.f);
  if (gb && i == 0) gb[o] = s0[o];
}

// Workspace regions: two big ping-pong buffers for the chain intermediates and two
// mode-space buffers (x_hat / y_hat in forward, g_yhat / g_xhat in backward).
struct Regions {
  float* big[2];
  float* mode[2];
  float* wprep[2];  // tf32-rounded (transposed) bypass weights of the tcgen05 stages
  float* lift;      // 4 x 128 floats: folded factors / partial reductions of a pointwise lifting in front of the block
};

// x_hat[B, C, J] <- x[B, C, *S]. Intermediates live in r.big[]; src/dst must not alias them.
// r2c_done: the first stage was already computed into the chain's first buffer (fused into another pass over x).
static int forward_chain(const pdeqx_spectral_plan* p, const float* x, int C, bool adjoint_of_inverse, float* dst,
                         const Regions& r, cudaStream_t s, bool r2c_done = false) {
  const pdeqx_spectral_desc& d = p->desc;
  const int D = d.ndim;
  const int mD = p->mlast;
  const int64_t BC = (int64_t)d.batch * C;
  const float* tab_last = adjoint_of_inverse ? p->t_r2c_bwd : p->t_r2c_fwd;
  float* o = (D == 1) ? dst : r.big[0];
  if (!r2c_done) PDEQX_TRY(stage_r2c(p, x, tab_last, BC * p->lead, o, s));  // else: the caller already filled `o`
  const float* cur = o;
  int nb = 1;
  // axes D-2 .. 0; at axis a the leading axes a' < a are still full-size, trailing ones are truncated
  for (int a = D - 2; a >= 0; --a) {
    int64_t outer = BC, inner = mD;
    for (int b = 0; b < a; ++b) outer *= d.spatial[b];
    for (int b = a + 1; b < D - 1; ++b) inner *= p->n2[b];
    float* dstage = (a == 0) ? dst : r.big[nb];
    const float2* tab = adjoint_of_inverse ? p->t_inv_adj[a] : p->t_fwd[a];
    PDEQX_TRY(stage_c2c(p, a, adjoint_of_inverse ? 2 : 0, tab, cur, dstage, outer, d.spatial[a], p->n2[a], inner, s));
    cur = dstage;
    nb ^= 1;
  }
  return PDEQX_OK;
}

// z[B*C*lead, m_D] <- y_hat[B, C, J] (leading-axis inverse stages, D > 1). Intermediates in r.big[0] only.
static int inverse_chain(const pdeqx_spectral_plan* p, const float* yhat, int C, bool adjoint_of_forward, float* dst,
                         const Regions& r, cudaStream_t s) {
  const pdeqx_spectral_desc& d = p->desc;
  const int D = d.ndim;
  const int mD = p->mlast;
  const int64_t BC = (int64_t)d.batch * C;
  const float* cur = yhat;
  for (int a = 0; a < D - 1; ++a) {
    int64_t outer = BC, inner = mD;
    for (int b = 0; b < a; ++b) outer *= d.spatial[b];
    for (int b = a + 1; b < D - 1; ++b) inner *= p->n2[b];
    float* dstage = (a == D - 2) ? dst : r.big[0];
    const float2* tab = adjoint_of_forward ? p->t_fwd_adj[a] : p->t_inv[a];
    PDEQX_TRY(stage_c2c(p, a, adjoint_of_forward ? 3 : 1, tab, cur, dstage, outer, p->n2[a], d.spatial[a], inner, s));
    cur = dstage;
  }
  return PDEQX_OK;
}

// Samples per chunk of the L2-blocked front of the backward pass (0 = unchunked; experiment knob PDEQX_BWD_CHUNK).
static int bwd_chunk_samples(const pdeqx_spectral_plan* p, int64_t n_pix) {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("PDEQX_BWD_CHUNK");
    env = e ? atoi(e) : -1;
  }
  (void)n_pix;
  // Measured on B200 (configs[3], B = 64): 1.71 ms unchunked vs 1.87 / 2.08 / 2.48 / 3.13 ms at 16 / 8 / 4 / 2 samples per
  // chunk -- the fixed cost of three persistent-kernel launches per chunk outweighs the L2 hits, so the default is off.
  return env > 0 && env < p->desc.batch ? env : 0;
}

static int check_ws(const pdeqx_spectral_plan* p, void* ws, size_t ws_bytes) {
  if (!ws || ws_bytes < p->ws_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", p->ws_bytes, ws_bytes);
    return PDEQX_ERR_WORKSPACE;
  }
  return PDEQX_OK;
}

static Regions make_regions(const pdeqx_spectral_plan* p, void* ws) {
  uintptr_t base = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  Regions r;
  r.big[0] = (float*)base;
  r.big[1] = r.big[0] + p->stage_floats;
  r.mode[0] = r.big[1] + p->stage_floats;
  r.mode[1] = r.mode[0] + p->mode_floats;
  const int64_t cmax = p->desc.c_in > p->desc.c_out ? p->desc.c_in : p->desc.c_out;
  r.wprep[0] = r.mode[1] + p->mode_floats;
  r.wprep[1] = r.wprep[0] + cmax * cmax;
  r.lift = r.wprep[1] + cmax * cmax;
  return r;
}

}  // namespace pdeqx

// lift_u != null: the block's input is the pointwise lifting lift_w[i] * lift_u[b, pixel] + lift_b[i] of a single-channel
// field and is never materialised (x is ignored): its first DFT stage is the lifting of the field's first stage, and its
// bypass term is rank one, folded into the last kernel's epilogue
// proj_out != null: the block's output only feeds a pointwise C -> 1 projection out[b,pixel] = sum_o proj_w[o] y[b,o,pixel] +
// proj_b; it is reduced in the last kernel's epilogue and y is never written (y may be null)
static int spectral_block_fwd_impl(const pdeqx_spectral_plan* p, const float* x, const float* lift_u, const float* lift_w,
                                   const float* lift_b, const float* proj_w, const float* proj_b, float* proj_out,
                                   const float* w_re,
                                   const float* w_im, const float* bypass_w, const float* bypass_b,
                                   int32_t act, float* y, float* save_xhat, float* save_z, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(p && (x || lift_u) && w_re && w_im && (y || proj_out), "null argument");
  PDEQX_CHECK_ARG(act >= 0 && act <= 2, "unknown activation %d", act);
  PDEQX_CHECK_ARG(bypass_w || !bypass_b, "bypass_b without bypass_w");
  PDEQX_TRY(check_ws(p, workspace, workspace_bytes));
  cudaStream_t s = (cudaStream_t)stream;
  const pdeqx_spectral_desc& d = p->desc;
  const int D = d.ndim;
  const int sD = d.spatial[D - 1];
  const int64_t n_pix = p->lead * sD;
  Regions r = make_regions(p, workspace);

  // forward transform -> x_hat [B, c_in, J]
  float* xhat = save_xhat ? save_xhat : r.mode[0];
  const bool lifted = lift_u != nullptr;
  if (lifted) {
    PDEQX_CHECK_ARG(lift_w && bypass_w && D >= 2 && d.c_in <= 128 && d.c_out <= 128 &&
                        tc_slab_available(p, true) && tc_r2c_available(p) && p->mlast == 16,
                    "the lifted forward pass needs the tcgen05 path (pdeqx_spectral_block_rank1_ok) and a bypass");
    // first stage of the single-channel field, then its lifting (the DFT is linear; the bias only feeds column k = 0)
    float* u1 = r.big[1];
    PDEQX_TRY(stage_r2c(p, lift_u, p->t_r2c_fwd, (int64_t)d.batch * p->lead, u1, s));
    const int64_t total4 = (int64_t)d.batch * d.c_in * p->lead * 8;
    const int blocks = (int)(ceil_div(total4, 256) < 148 * 16 ? ceil_div(total4, 256) : 148 * 16);
    PDEQX_CUDA(launch_pdl(lift_x1_kernel, dim3((unsigned)(blocks)), dim3(256), 0, s, (const float4*)u1, lift_w, lift_b, (float)sD, d.c_in, p->lead, total4, 1, (float4*)r.big[0]));
    PDEQX_LAUNCHED();
    PDEQX_TRY(forward_chain(p, nullptr, d.c_in, false, xhat, r, s, /*r2c_done=*/true));
  } else {
    PDEQX_TRY(forward_chain(p, x, d.c_in, false, xhat, r, s));
  }
  // channel mix -> y_hat [B, c_out, J] (for D == 1 this already is z)
  float* yhat = (D == 1 && save_z) ? save_z : r.mode[1];
  PDEQX_TRY(mode_mix(p, xhat, w_re, w_im, /*transposed_conj=*/false, yhat, s));
  // leading-axis inverse stages -> z
  const float* z = yhat;
  if (D > 1) {
    float* zbuf = save_z ? save_z : r.big[1];
    PDEQX_TRY(inverse_chain(p, yhat, d.c_out, false, zbuf, r, s));
    z = zbuf;
  }
  // last stage: c2r (+ bypass + bias + activation)
  const int64_t rows = (int64_t)d.batch * d.c_out * p->lead;
  const bool projected = proj_out != nullptr;
  if (projected) {
    PDEQX_CHECK_ARG(proj_w && bypass_w && tc_slab_available(p, true),
                    "the projected forward pass needs the tcgen05 path (pdeqx_spectral_block_rank1_ok) and a bypass");
    const int64_t n1 = (int64_t)d.batch * n_pix;
    PDEQX_CUDA(launch_pdl(fill_kernel, dim3((unsigned)((int)(ceil_div(n1, 1024) < 148 * 8 ? ceil_div(n1, 1024) : 148 * 8))), dim3(256), 0, s, proj_out, n1, proj_b));
    PDEQX_LAUNCHED();
  }
  if (lifted) {
    float* fa = r.lift;       // a[o]: per-channel factor of the rank-one bypass
    float* fc = r.lift + 128;  // c[o]: its constant part (+ bypass bias)
    PDEQX_CUDA(launch_pdl(lift_fold_bypass_kernel, dim3((unsigned)((d.c_out + 127) / 128)), dim3(128), 0, s, bypass_w, bypass_b, lift_w, lift_b, d.c_out, d.c_in, fa, fc));
    PDEQX_LAUNCHED();
    TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
    return tc_slab_ex(p, r.wprep[0], nullptr, z, /*table_sel=*/0, nullptr, false, fc, act, /*mode=*/projected ? 4 : 0, nullptr, d.c_in, d.c_out,
                      projected ? proj_out : y, s, projected ? proj_w : nullptr, nullptr, nullptr, lift_u, fa);
  }
  if (projected) {
    TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
    return tc_slab_ex(p, r.wprep[0], x, z, /*table_sel=*/0, bypass_w, false, bypass_b, act, /*mode=*/4, nullptr, d.c_in, d.c_out, proj_out, s, proj_w);
  }
  if (tc_slab_available(p, bypass_w != nullptr)) {
    TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
    return tc_slab(p, r.wprep[0], x, z, /*table_sel=*/0, bypass_w, /*bypass_transposed=*/false, bypass_b, act, y, s);
  }
  if (!bypass_w) return stage_c2r(p, z, p->t_c2r_inv, rows, nullptr, act, y, s);
  PDEQX_TRY(stage_c2r(p, z, p->t_c2r_inv, rows, nullptr, PDEQX_ACT_IDENTITY, y, s));
  return pdeqx_pointwise_fwd(d.batch, d.c_in, d.c_out, n_pix, x, bypass_w, bypass_b, /*add=*/y, act, y, stream);
}

// gy_w != null: the cotangent is rank one, gy[b,o,pixel] = gy_w[o] * gy[b,pixel] (gy then holds the single-channel factor)
extern "C" int pdeqx_spectral_block_fwd(const pdeqx_spectral_plan* p, const float* x, const float* w_re,
                                        const float* w_im, const float* bypass_w, const float* bypass_b,
                                        int32_t act, float* y, float* save_xhat, float* save_z, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  PDEQX_CHECK_ARG(x, "null argument");
  return spectral_block_fwd_impl(p, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, w_re, w_im, bypass_w, bypass_b, act, y,
                                 save_xhat, save_z, workspace, workspace_bytes, stream);
}

extern "C" int pdeqx_spectral_block_fwd_ex(const pdeqx_spectral_plan* p, const float* x, const float* lift_u,
                                           const float* lift_w, const float* lift_b, const float* proj_w,
                                           const float* proj_b, float* proj_out, const float* w_re, const float* w_im,
                                           const float* bypass_w, const float* bypass_b, int32_t act, float* y,
                                           float* save_xhat, float* save_z, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  PDEQX_CHECK_ARG((x != nullptr) != (lift_u != nullptr), "give either x or the lifted field");
  PDEQX_CHECK_ARG(!lift_u || lift_w, "null argument");
  return spectral_block_fwd_impl(p, x, lift_u, lift_w, lift_b, proj_w, proj_b, proj_out, w_re, w_im, bypass_w, bypass_b, act, y,
                                 save_xhat, save_z, workspace, workspace_bytes, stream);
}

// lift_u != null: gx is not produced; the block's input came from the single-channel field lift_u through a 1 -> c_in
// pointwise lifting, whose gradients lift_gw[i] = sum gx[b,i,pixel] lift_u[b,pixel], lift_gb[i] = sum gx[b,i,pixel] are
// accumulated by the last kernel instead
static int spectral_block_bwd_impl(const pdeqx_spectral_plan* p, const pdeqx_spectral_bwd_args& a) {
  const float *x = a.x, *gy = a.gy, *gy_w = a.gy_w, *lift_u = a.lift_u, *lift_w = a.lift_w, *lift_b = a.lift_b;
  float *lift_gw = a.lift_gw, *lift_gb = a.lift_gb, *proj_gw = a.proj_gw;
  const float *w_re = a.w_re, *w_im = a.w_im, *bypass_w = a.bypass_w, *bypass_b = a.bypass_b;
  const int32_t act = a.activation;
  const float *saved_xhat = a.saved_xhat, *saved_z = a.saved_z;
  float *gx = a.gx, *gw_re = a.gw_re, *gw_im = a.gw_im, *g_bypass_w = a.g_bypass_w, *g_bypass_b = a.g_bypass_b;
  float* scratch_full = a.scratch_full;
  void* workspace = a.workspace;
  const size_t workspace_bytes = a.workspace_bytes;
  void* stream = a.stream;
  const bool chained = a.up_plan != nullptr;      // the data-gradient pass also forms the upstream block's pre-activation cotangent
  const bool gy_is_gpre = a.gy_is_gpre != 0;      // the downstream block's chained pass already applied act'(pre) of THIS block
  PDEQX_CHECK_ARG(p && (x || (lift_u && lift_w)) && gy && w_re && w_im && saved_xhat && gw_re && gw_im, "null argument");
  PDEQX_CHECK_ARG(act >= 0 && act <= 2, "unknown activation %d", act);
  PDEQX_CHECK_ARG(!bypass_w || g_bypass_w, "g_bypass_w is required when bypass_w is given");
  // x == null: the block's input lift_w (x) lift_u + lift_b was never materialised (forward ran through
  // pdeqx_spectral_block_fwd_lifted); every place that reads x below works from the single-channel field instead
  const bool lifted = x == nullptr;
  PDEQX_TRY(check_ws(p, workspace, workspace_bytes));
  cudaStream_t s = (cudaStream_t)stream;
  const pdeqx_spectral_desc& d = p->desc;
  const int D = d.ndim;
  const int sD = d.spatial[D - 1];
  const int64_t n_pix = p->lead * sD;
  const int64_t rows_o = (int64_t)d.batch * d.c_out * p->lead;
  const int64_t rows_i = (int64_t)d.batch * d.c_in * p->lead;
  Regions r = make_regions(p, workspace);

  const bool fast = tc_slab_available(p, bypass_w != nullptr);
  const float* gpre = gy;
  bool gx_has_bypass = false;
  PDEQX_CHECK_ARG(!gy_is_gpre || (!gy_w && !proj_gw), "gy_is_gpre excludes the rank-one cotangent arguments");
  if (chained) {
    PDEQX_CHECK_ARG(!gx && !lift_u, "a chained reverse pass writes up_gpre instead of gx");
    PDEQX_CHECK_ARG(fast && bypass_w && tc_chain_available(p, a.up_plan) && a.up_saved_z && a.up_gpre && a.up_bypass_w &&
                        (a.up_x || (a.up_lift_u && a.up_lift_w)) &&
                        (a.up_activation == PDEQX_ACT_RELU || a.up_activation == PDEQX_ACT_GELU_TANH),
                    "the chained reverse pass does not qualify (pdeqx_spectral_block_chain_ok)");
  }
  const bool fuse_bypass_gx = fast && bypass_w && gx;
  bool r2c_done = false;
  const int chunk = bwd_chunk_samples(p, n_pix);
  PDEQX_CHECK_ARG(!lifted || (fast && act != PDEQX_ACT_IDENTITY && bypass_w && lift_gw),
                  "the lifted reverse pass needs the tcgen05 slab kernel, a fused activation and a bypass");
  PDEQX_CHECK_ARG(!proj_gw || (gy_w && d.c_out <= 64), "proj_gw needs the rank-one cotangent (gy_w) and c_out <= 64");
  if (proj_gw) PDEQX_TRY(zero_small(proj_gw, d.c_out, nullptr, 0, s));
  PDEQX_CHECK_ARG(!gy_w || (fast && act != PDEQX_ACT_IDENTITY),
                  "a rank-1 cotangent needs the tcgen05 slab kernel and a fused activation (pdeqx_spectral_block_rank1_ok)");
  if (!gy_is_gpre && !gy_w && !lifted && fast && chunk > 0 && act != PDEQX_ACT_IDENTITY && bypass_w && tc_wgrad_available(p, n_pix) && tc_r2c_available(p)) {
    // Steps 1-3a in sample chunks small enough for the 126 MB L2: the cotangent chunk written by the slab kernel is
    // still L2-resident when the bypass weight gradient and the first DFT stage read it back, and so is most of x.
    PDEQX_CHECK_ARG(saved_z && scratch_full, "saved_z and scratch_full are required when an activation is fused");
    const int mD = p->mlast;
    float* x1 = (D == 1) ? r.mode[0] : r.big[0];
    float* gb = bypass_b ? g_bypass_b : nullptr;
    PDEQX_TRY(zero_small(g_bypass_w, (int64_t)d.c_in * d.c_out, gb, d.c_out, s));
    for (int b0 = 0; b0 < d.batch; b0 += chunk) {
      pdeqx_spectral_plan pc = *p;  // shallow copy: same tables, fewer samples
      pc.desc.batch = d.batch - b0 < chunk ? d.batch - b0 : chunk;
      const int64_t off_o = (int64_t)b0 * d.c_out * n_pix, off_i = (int64_t)b0 * d.c_in * n_pix;
      const int64_t off_z = (int64_t)b0 * d.c_out * p->lead * 2 * mD;
      {
        TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
        PDEQX_TRY(tc_slab_ex(&pc, r.wprep[0], x + off_i, saved_z + off_z, 0, bypass_w, false, bypass_b, act, /*mode=*/1, gy + off_o,
                             d.c_in, d.c_out, scratch_full + off_o, s));
      }
      {
        TimedScope timed(PDEQX_K_POINTWISE_WGRAD, s);
        PDEQX_TRY(tc_wgrad(&pc, scratch_full + off_o, x + off_i, n_pix, g_bypass_w, gb, s));
      }
      PDEQX_TRY(stage_r2c(p, scratch_full + off_o, p->t_r2c_bwd, (int64_t)pc.desc.batch * d.c_out * p->lead, x1 + off_z, s));
    }
    gpre = scratch_full;
    r2c_done = true;
  } else {
  // 1. cotangent of the pre-activation (pre-activation recomputed from the saved z: nothing full-size is kept)
  if (act != PDEQX_ACT_IDENTITY && !gy_is_gpre) {
    PDEQX_CHECK_ARG(saved_z && scratch_full, "saved_z and scratch_full are required when an activation is fused");
    if (lifted) {
      float* fa = r.lift;
      float* fc = r.lift + 128;
      PDEQX_CUDA(launch_pdl(lift_fold_bypass_kernel, dim3((unsigned)((d.c_out + 127) / 128)), dim3(128), 0, s, bypass_w, bypass_b, lift_w, lift_b, d.c_out, d.c_in, fa, fc));
      PDEQX_LAUNCHED();
      TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
      PDEQX_TRY(tc_slab_ex(p, r.wprep[0], nullptr, saved_z, 0, nullptr, false, fc, act, /*mode=*/gy_w ? (proj_gw ? 5 : 2) : 1, gy, d.c_in, d.c_out,
                           scratch_full, s, gy_w, proj_gw, nullptr, lift_u, fa));
    } else if (fast) {
      TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
      // mode 2: the cotangent tile is formed in the epilogue from its two factors instead of being read from HBM
      // mode 5: ... and the projection's weight gradient is accumulated from the recomputed activation (the block's output
      // was never stored by the projected forward pass)
      PDEQX_TRY(tc_slab_ex(p, r.wprep[0], x, saved_z, 0, bypass_w, false, bypass_b, act, /*mode=*/gy_w ? (proj_gw ? 5 : 2) : 1, gy, d.c_in, d.c_out,
                           scratch_full, s, gy_w, proj_gw));
    } else {
      PDEQX_TRY(stage_c2r(p, saved_z, p->t_c2r_inv, rows_o, nullptr, PDEQX_ACT_IDENTITY, scratch_full, s));
      if (bypass_w)
        PDEQX_TRY(pdeqx_pointwise_fwd(d.batch, d.c_in, d.c_out, n_pix, x, bypass_w, bypass_b, scratch_full,
                                      PDEQX_ACT_IDENTITY, scratch_full, stream));
      PDEQX_TRY(pdeqx_activation_bwd((int64_t)d.batch * d.c_out * n_pix, scratch_full, gy, act, scratch_full, stream));
    }
    gpre = scratch_full;
  }
  // 2. bypass grads (and the bypass part of gx; on the tcgen05 path that part is fused into step 7)
  if (bypass_w && lifted) {
    // gW_byp[o,i] = sum g[o,p] x[i,p] with x = lift_w (x) u + lift_b: two reductions over g (one pass), then an outer product
    float* s1 = r.lift + 256;
    float* s0 = r.lift + 384;
    TimedScope timed(PDEQX_K_POINTWISE_WGRAD, s);
    if (fast && tc_wgrad_available(p, n_pix) && fuse_r2c_mode() == 1 && tc_wgrad_r2c_available(p)) {
      // the two reductions are the bypass weight gradient against a ONE-channel input: the tcgen05 weight-gradient kernel
      // takes the field as its x operand, and its pass over gpre also applies the first DFT stage (no second read of gpre)
      PDEQX_TRY(zero_small(s1, d.c_out, s0, d.c_out, s));
      PDEQX_TRY(tc_wgrad(p, gpre, lift_u, n_pix, s1, s0, s, (D == 1) ? r.mode[0] : r.big[0], /*rank1=*/true));
      r2c_done = true;
    } else {
      PDEQX_TRY(pdeqx_pointwise_bwd(d.batch, 1, d.c_out, n_pix, lift_u, lift_w, gpre, nullptr, s1, s0, stream));
    }
    PDEQX_CUDA(launch_pdl(lift_unfold_wgrad_kernel, dim3((unsigned)((d.c_out * d.c_in + 255) / 256)), dim3(256), 0, s, s1, s0, lift_w, lift_b, d.c_out, d.c_in, g_bypass_w,
                                                                         bypass_b ? g_bypass_b : nullptr));
    PDEQX_LAUNCHED();
  } else if (bypass_w) {
    if (fast && tc_wgrad_available(p, n_pix)) {
      TimedScope timed(PDEQX_K_POINTWISE_WGRAD, s);
      float* gb = bypass_b ? g_bypass_b : nullptr;
      PDEQX_TRY(zero_small(g_bypass_w, (int64_t)d.c_in * d.c_out, gb, d.c_out, s));
      // step 3a rides along: the kernel that streams gpre for the weight gradient also applies the first DFT stage
      float* x1 = nullptr;
      if (fuse_r2c_mode() == 1 && tc_wgrad_r2c_available(p)) {
        x1 = (D == 1) ? r.mode[0] : r.big[0];
        r2c_done = true;
      }
      PDEQX_TRY(tc_wgrad(p, gpre, x, n_pix, g_bypass_w, gb, s, x1));
    } else {
      PDEQX_TRY(pdeqx_pointwise_bwd(d.batch, d.c_in, d.c_out, n_pix, x, bypass_w, gpre, fuse_bypass_gx ? nullptr : gx,
                                    g_bypass_w, bypass_b ? g_bypass_b : nullptr, stream));
      gx_has_bypass = gx != nullptr && !fuse_bypass_gx;
    }
  }
  }
  // 3. g_yhat = adjoint of the inverse transform applied to gpre
  float* ghat = r.mode[0];
  PDEQX_TRY(forward_chain(p, gpre, d.c_out, /*adjoint_of_inverse=*/true, ghat, r, s, r2c_done));
  // 4. weight gradient (batch-reduced inside the kernel)
  PDEQX_TRY(mode_mix_dw(p, ghat, saved_xhat, gw_re, gw_im, s));
  PDEQX_CHECK_ARG(!lift_u || (fast && bypass_w), "the fused lifting gradients need the tcgen05 slab kernel (pdeqx_spectral_block_rank1_ok)");
  if (!gx && !lift_u && !chained) return PDEQX_OK;
  // 5. g_xhat = W^H g_yhat
  float* gxhat = r.mode[1];
  PDEQX_TRY(mode_mix(p, ghat, w_re, w_im, /*transposed_conj=*/true, gxhat, s));
  // 6. adjoint of the leading forward stages
  const float* gx1 = gxhat;
  if (D > 1) {
    float* t = r.big[1];
    PDEQX_TRY(inverse_chain(p, gxhat, d.c_in, /*adjoint_of_forward=*/true, t, r, s));
    gx1 = t;
  }
  // 7. last stage back to real space, plus the bypass contribution W^T gpre
  if (chained) {
    // ... and, in the same pass, times act'(pre) of the block in front of this one: this block's input gradient only ever
    // exists as an accumulator tile; what is written is the upstream block's pre-activation cotangent
    TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
    if (!a.up_x) {
      float* fa = r.lift;
      float* fc = r.lift + 128;
      PDEQX_CUDA(launch_pdl(lift_fold_bypass_kernel, dim3((unsigned)((d.c_in + 127) / 128)), dim3(128), 0, s, a.up_bypass_w, a.up_bypass_b, a.up_lift_w, a.up_lift_b, d.c_in,
                                                                  a.up_plan->desc.c_in, fa, fc));
      PDEQX_LAUNCHED();
      return tc_chain(p, r.wprep[0], r.wprep[1], gpre, gx1, bypass_w, a.up_plan, nullptr, a.up_saved_z, nullptr, fc, a.up_activation, a.up_gpre, s, a.up_lift_u, fa);
    }
    return tc_chain(p, r.wprep[0], r.wprep[1], gpre, gx1, bypass_w, a.up_plan, a.up_x, a.up_saved_z, a.up_bypass_w, a.up_bypass_b, a.up_activation, a.up_gpre, s);
  }
  if (fast && lift_u) {
    TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
    PDEQX_TRY(zero_small(lift_gw, d.c_in, lift_gb, d.c_in, s));
    return tc_slab_ex(p, r.wprep[0], gpre, gx1, /*table_sel=*/1, bypass_w, /*bypass_transposed=*/true, nullptr, PDEQX_ACT_IDENTITY, /*mode=*/3,
                      lift_u, d.c_out, d.c_in, nullptr, s, nullptr, lift_gw, lift_gb);
  }
  if (fast) {
    TimedScope timed(PDEQX_K_SPECTRAL_SLAB, s);
    return tc_slab(p, r.wprep[0], gpre, gx1, /*table_sel=*/1, bypass_w, /*bypass_transposed=*/true, nullptr, PDEQX_ACT_IDENTITY, gx, s);
  }
  return stage_c2r(p, gx1, p->t_c2r_bwd, rows_i, gx_has_bypass ? gx : nullptr, PDEQX_ACT_IDENTITY, gx, s);
}

extern "C" int pdeqx_spectral_block_bwd_args(const pdeqx_spectral_plan* p, const pdeqx_spectral_bwd_args* args) {
  PDEQX_CHECK_ARG(args, "null argument");
  PDEQX_CHECK_ARG(!args->lift_u || (args->lift_gw && !args->gx), "lift_u needs lift_gw and replaces gx");
  return spectral_block_bwd_impl(p, *args);
}

extern "C" int pdeqx_spectral_block_chain_ok(const pdeqx_spectral_plan* p, const pdeqx_spectral_plan* up, int32_t up_activation) {
  return (up_activation == PDEQX_ACT_RELU || up_activation == PDEQX_ACT_GELU_TANH) && tc_chain_available(p, up) ? 1 : 0;
}

extern "C" int pdeqx_spectral_block_bwd(const pdeqx_spectral_plan* p, const float* x, const float* gy,
                                        const float* w_re, const float* w_im, const float* bypass_w,
                                        const float* bypass_b, int32_t act, const float* saved_xhat,
                                        const float* saved_z, float* gx, float* gw_re, float* gw_im,
                                        float* g_bypass_w, float* g_bypass_b, float* scratch_full, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  pdeqx_spectral_bwd_args a{};
  a.x = x; a.gy = gy; a.w_re = w_re; a.w_im = w_im; a.bypass_w = bypass_w; a.bypass_b = bypass_b; a.activation = act;
  a.saved_xhat = saved_xhat; a.saved_z = saved_z; a.gx = gx; a.gw_re = gw_re; a.gw_im = gw_im;
  a.g_bypass_w = g_bypass_w; a.g_bypass_b = g_bypass_b; a.scratch_full = scratch_full;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes; a.stream = stream;
  return spectral_block_bwd_impl(p, a);
}

extern "C" int pdeqx_spectral_block_rank1_ok(const pdeqx_spectral_plan* p, int32_t has_bypass, int32_t act) {
  return p && act != PDEQX_ACT_IDENTITY && tc_slab_available(p, has_bypass != 0) ? 1 : 0;
}

extern "C" int pdeqx_spectral_block_bwd_ex(const pdeqx_spectral_plan* p, const float* x, const float* gy,
                                           const float* gy_w, const float* lift_u, float* lift_gw, float* lift_gb,
                                           const float* lift_w, const float* lift_b, float* proj_gw,
                                           const float* w_re, const float* w_im, const float* bypass_w,
                                           const float* bypass_b, int32_t act, const float* saved_xhat,
                                           const float* saved_z, float* gx, float* gw_re, float* gw_im,
                                           float* g_bypass_w, float* g_bypass_b, float* scratch_full, void* workspace,
                                           size_t workspace_bytes, void* stream) {
  pdeqx_spectral_bwd_args a{};
  a.x = x; a.gy = gy; a.gy_w = gy_w; a.lift_u = lift_u; a.lift_gw = lift_gw; a.lift_gb = lift_gb; a.lift_w = lift_w;
  a.lift_b = lift_b; a.proj_gw = proj_gw;
  a.w_re = w_re; a.w_im = w_im; a.bypass_w = bypass_w; a.bypass_b = bypass_b; a.activation = act;
  a.saved_xhat = saved_xhat; a.saved_z = saved_z; a.gx = gx; a.gw_re = gw_re; a.gw_im = gw_im;
  a.g_bypass_w = g_bypass_w; a.g_bypass_b = g_bypass_b; a.scratch_full = scratch_full;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes; a.stream = stream;
  return pdeqx_spectral_block_bwd_args(p, &a);
}
