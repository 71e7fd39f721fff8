// Internal: the spectral plan object and the stage entry points shared between
// spectral.cu (sequencing), modemix.cu (channel mix) and tc_spectral.cu (tcgen05 stages).
#pragma once
#include "common.cuh"

struct pdeqx_spectral_plan {
  pdeqx_spectral_desc desc;
  int G = 1;
  int64_t lead = 1;          // prod of the leading spatial sizes s_0..s_{D-2}
  int mlast = 1;             // mode count of the DATA layout on the contiguous axis (>= modes[D-1]; padded modes are zero)
  int n2[PDEQX_MAX_DIMS] = {1, 1, 1};  // retained rows of the data layout on each leading axis (>= 2 m_a)
  int64_t J = 1;             // retained (padded) modes per channel: prod(n2_a, a < D-1) * mlast
  int64_t stage_floats = 0;  // floats of one chain ping-pong buffer
  int64_t mode_floats = 0;   // floats of one mode-space buffer
  size_t ws_bytes = 0;
  int64_t small_floats = 0;  // per-call scratch of the tcgen05 stages inside the workspace (bypass weights, lifting factors)
  // last axis, real tables (fp32, generated in fp64)
  float* t_r2c_fwd = nullptr;  // [s_D, 2 m_D]  cos | -sin                (forward)
  float* t_c2r_inv = nullptr;  // [2 m_D, s_D]  c/s cos ; -c/s sin        (inverse, irfft weights c)
  float* t_r2c_bwd = nullptr;  // [s_D, 2 m_D]  = t_c2r_inv^T             (adjoint of inverse)
  float* t_c2r_bwd = nullptr;  // [2 m_D, s_D]  = t_r2c_fwd^T             (adjoint of forward)
  // leading axes, complex tables
  float2* t_fwd[PDEQX_MAX_DIMS] = {nullptr, nullptr, nullptr};      // [2m, s]
  float2* t_inv[PDEQX_MAX_DIMS] = {nullptr, nullptr, nullptr};      // [s, 2m]
  float2* t_inv_adj[PDEQX_MAX_DIMS] = {nullptr, nullptr, nullptr};  // [2m, s] = conj(t_inv)^T
  float2* t_fwd_adj[PDEQX_MAX_DIMS] = {nullptr, nullptr, nullptr};  // [s, 2m] = conj(t_fwd)^T
  // tcgen05 side (tc_spectral.cu); null when the shape does not qualify
  void* tc = nullptr;
  // tcgen05 leading-axis stages (tc_c2c.cu)
  void* tc_c2c = nullptr;
};

namespace pdeqx {

// out[b,u,p] = sum_v W(u,v,p) in[b,v,p]; transposed_conj: W(u,v) = conj(w[o=v, i=u]) (u over c_in), else
// W(u,v) = w[o=u, i=v] (u over c_out). in/out are complex-interleaved [B, C, J].
int mode_mix(const pdeqx_spectral_plan* p, const float* in, const float* w_re, const float* w_im,
             bool transposed_conj, float* out, cudaStream_t s);
// gw[g,o,i,loc] = sum_b ghat[b,o,p] conj(xhat[b,i,p])
int mode_mix_dw(const pdeqx_spectral_plan* p, const float* ghat, const float* xhat, float* gw_re, float* gw_im,
                cudaStream_t s);

// the same two on tcgen05 (tc_modemix.cu; TF32 precision)
bool tc_mode_gemm_available(const pdeqx_spectral_plan* p);
int tc_mode_mix(const pdeqx_spectral_plan* p, const float* in, const float* w_re, const float* w_im, bool transposed_conj,
                float* out, cudaStream_t s);
int tc_mode_mix_dw(const pdeqx_spectral_plan* p, const float* ghat, const float* xhat, float* gw_re, float* gw_im,
                   cudaStream_t s);

// tcgen05 stages
int tc_plan_init(pdeqx_spectral_plan* p);
void tc_plan_free(pdeqx_spectral_plan* p);
bool tc_r2c_available(const pdeqx_spectral_plan* p);
int tc_stage_mask(const pdeqx_spectral_plan* p);
int tc_r2c(const pdeqx_spectral_plan* p, const float* x, bool adjoint_table, int64_t rows, float* out, cudaStream_t s);
bool tc_slab_available(const pdeqx_spectral_plan* p, bool has_bypass);
// y = act( c2r(z) [table_sel: 0 = inverse, 1 = adjoint-of-forward] + W(.)x + b ); w_prep: c_in * c_out floats of caller scratch
int tc_slab(const pdeqx_spectral_plan* p, float* w_prep, const float* x, const float* z, int table_sel, const float* bypass_w,
            bool bypass_transposed, const float* bypass_b, int act, float* y, cudaStream_t s);

// leading-axis complex DFT stages on tcgen05 (tc_c2c.cu). kind: 0 forward, 1 inverse, 2 adjoint-of-inverse, 3 adjoint-of-forward
int tc_c2c_plan_init(pdeqx_spectral_plan* p);
void tc_c2c_plan_free(pdeqx_spectral_plan* p);
bool tc_c2c_available(const pdeqx_spectral_plan* p, int axis, int kind, int64_t inner_complex);
// round_out: the result feeds another tcgen05 (tf32) stage -- round it to tf32 (nearest) here instead of letting that
// stage truncate the operand
int tc_c2c(const pdeqx_spectral_plan* p, int axis, int kind, const float* in, float* out, int64_t outer, int64_t inner_complex,
           cudaStream_t s, bool round_out = false);
// general form: explicit channel counts of THIS call, epilogue mode 1 = aux * act'(.) (cotangent of the pre-activation),
// mode 2 = the same with a rank-1 cotangent aux_w[o] * aux[b, pixel] (behind a C -> 1 pointwise projection),
// byp_u / byp_a (modes 0-2, bypass_w == null): rank-one bypass byp_a[o] * byp_u[b, pixel] added before the activation;
// mode 3 = nothing stored: red_w[o] += sum result * aux[b, pixel], red_b[o] += sum result (in front of a 1 -> C lifting)
int tc_slab_ex(const pdeqx_spectral_plan* p, float* w_prep, const float* x, const float* z, int table_sel, const float* bypass_w,
               bool bypass_transposed, const float* bypass_b, int act, int mode, const float* aux, int c_in, int c_out,
               float* y, cudaStream_t s, const float* aux_w = nullptr, float* red_w = nullptr, float* red_b = nullptr,
               const float* byp_u = nullptr, const float* byp_a = nullptr);
// the data-gradient pass of block `p` fused with the pre-activation-cotangent pass of the block `up` in front of it
bool tc_chain_available(const pdeqx_spectral_plan* p, const pdeqx_spectral_plan* up);
int tc_chain(const pdeqx_spectral_plan* p, float* w_prep, float* w_prep2, const float* g, const float* z1, const float* w_k, const pdeqx_spectral_plan* up,
             const float* x_up, const float* z_up, const float* w_up, const float* bias_up, int act, float* out, cudaStream_t s,
             const float* byp_u = nullptr, const float* byp_a = nullptr);
bool tc_wgrad_available(const pdeqx_spectral_plan* p, int64_t n_pix);
// gw[c_out,c_in] += sum g x^T, gb[c_out] += sum g (outputs zeroed by the caller); rank1: x is a single-channel field
// [B, 1, pixels] (the input of a 1 -> C lifting that was folded into the block) and gw is [c_out]
int tc_wgrad(const pdeqx_spectral_plan* p, const float* g, const float* x, int64_t n_pix, float* gw, float* gb,
             cudaStream_t s, float* x1_out = nullptr, bool rank1 = false);
bool tc_wgrad_r2c_available(const pdeqx_spectral_plan* p);

}  // namespace pdeqx
