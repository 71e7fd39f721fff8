/* pdeqx_b200.h -- C ABI of libpdeqx_b200.so (sm_100a).
 *
 * Drop-in boundary for pdequinox's data-parallel hot path. The reference has no
 * native interface: its boundary is the Python eqx.Module call on one sample
 * under jax.vmap. These entry points are what an XLA FFI
 * custom call (jax.ffi.ffi_call inside jax.custom_vjp) binds for each of those
 * calls; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every tensor argument is a DEVICE pointer to contiguous float32, channels
 *     first, last spatial axis contiguous: x[B, C, *S] (the reference's
 *     single-sample (C,*S) convention with the vmap axis in front);
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as
 *     void*): no synchronisation, no allocation; scratch memory is passed in;
 *   - return value 0 = ok, negative = error (PDEQX_ERR_*); the message of the
 *     last error on the calling thread is pdeqx_last_error();
 *   - the library keeps no mutable state between calls: plans are immutable after
 *     creation (twiddle tables only) and may be shared by threads / streams of
 *     their device; every scratch buffer, including the small per-call ones
 *     (re-laid-out weights, partial sums), is carved from the workspace the
 *     caller passes, so calls are re-entrant as long as concurrent calls bring
 *     different workspaces. What IS process-wide: the per-device "kernel
 *     attribute configured" / "cluster launch unsupported" capability caches,
 *     the launch counter, the timing recorder (pdeqx_timing_enable) and the
 *     test switch pdeqx_set_fast_paths -- none of them changes results.
 */
#ifndef PDEQX_B200_H_
#define PDEQX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEQX_OK 0
#define PDEQX_ERR_INVALID_ARGUMENT (-1)
#define PDEQX_ERR_UNSUPPORTED (-2)
#define PDEQX_ERR_CUDA (-3)
#define PDEQX_ERR_WORKSPACE (-4)

#define PDEQX_MAX_DIMS 3

/* activation fused into an epilogue */
enum pdeqx_activation {
  PDEQX_ACT_IDENTITY = 0,
  PDEQX_ACT_RELU = 1,     /* jax.nn.relu  (ClassicResNet/DilatedResNet/ClassicUNet default) */
  PDEQX_ACT_GELU_TANH = 2 /* jax.nn.gelu(approximate=True) (ClassicFNO default, _classic_fno.py:20) */
};

/* halo construction; PhysicsConv maps periodic->WRAP, dirichlet->ZEROS,
 * neumann->REFLECT (pdequinox/conv/_physics_conv.py:92-101); EDGE is
 * MorePaddingConv's "replicate" (pdequinox/conv/_conv.py:198-200). */
enum pdeqx_boundary {
  PDEQX_PAD_ZEROS = 0,
  PDEQX_PAD_WRAP = 1,
  PDEQX_PAD_REFLECT = 2,
  PDEQX_PAD_EDGE = 3
};

/* arithmetic of the tensor-core stages. The CUDA-core stages are always fp32. */
enum pdeqx_precision {
  PDEQX_PREC_FP32 = 0,  /* fp32 parity mode: FFMA or 3xTF32 split; rel-L2 <= 1e-5 */
  PDEQX_PREC_TF32 = 1   /* single-pass TF32 on tcgen05; rel-L2 <= 2e-3 */
};

const char* pdeqx_last_error(void);
/* "major.minor.patch (sm_100a)" */
const char* pdeqx_version(void);
/* number of kernels this library has launched on the calling process since load
 * (bench.py's gpu_launches claim is read from here). */
uint64_t pdeqx_launch_count(void);
/* 1 if the tcgen05/TMA fast paths can run on the current device (sm_100). */
int pdeqx_has_tcgen05(void);
/* ordinal of the calling thread's current CUDA device (-1 on error): lets a binding key per-device caches (plans)
 * without linking the CUDA runtime itself. */
int pdeqx_current_device(void);
/* force-disable (0) / enable (1) the tcgen05 fast paths (tests compare both). */
void pdeqx_set_fast_paths(int enabled);

/* Per-kernel device timing for bench.py's roofline line: while enabled, the launchers of the tagged
 * kernels bracket each launch with a cudaEvent pair on the launching stream. pdeqx_timing_read
 * synchronises those events and returns the summed duration and the launch count of one tag. */
enum pdeqx_kernel_tag {
  PDEQX_K_SPECTRAL_SLAB = 1, /* last stage of the spectral block: c2r DFT + bypass 1x1 + bias + activation */
  PDEQX_K_SPECTRAL_R2C = 2,  /* first stage: truncated real-to-complex DFT along the contiguous axis */
  PDEQX_K_SPECTRAL_MODES = 3,/* leading-axis DFT stages + per-mode channel mix (mode-space kernels) */
  PDEQX_K_POINTWISE_WGRAD = 4,
  PDEQX_K_CONV_FWD = 5,
  PDEQX_K_CONV_BWD_DATA = 6,
  PDEQX_K_CONV_BWD_FILTER = 7,
  PDEQX_K_DP_EXCHANGE = 8,   /* pdeqx_dp_adam_step: reduce-scatter + Adam + all-gather (includes the wait for the slowest rank) */
  PDEQX_K_TAG_COUNT = 9
};
void pdeqx_timing_enable(int enabled); /* also clears what was recorded */
int pdeqx_timing_read(int32_t tag, double* total_ms, int64_t* launches);
/* same, plus the ALGORITHMIC bytes of those launches as stated by their launchers (every tensor the launch reads or
 * writes, counted once: SURVEY.md 8d) -- the numerator of bench.py's roofline.achieved. May be NULL. */
int pdeqx_timing_read_ex(int32_t tag, double* total_ms, int64_t* launches, double* algorithmic_bytes);

/* ------------------------------------------------------------------------- *
 * SpectralConv / ClassicSpectralBlock
 *   replaces pdequinox/conv/_spectral_conv.py:107-136 (spectral_conv_nd) and
 *   pdequinox/blocks/_classic_spectral_block.py:72-75 (block __call__)
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t ndim;                    /* 1..3 */
  int32_t batch;                   /* vmap axis */
  int32_t c_in, c_out;             /* call-time i / o of the weight (G, o, i, *modes) */
  int32_t spatial[PDEQX_MAX_DIMS]; /* S */
  int32_t modes[PDEQX_MAX_DIMS];   /* num_modes per axis */
  int32_t precision;               /* enum pdeqx_precision */
} pdeqx_spectral_desc;

typedef struct pdeqx_spectral_plan pdeqx_spectral_plan; /* opaque: twiddle tables on the device */

/* Builds twiddle tables in float64 on the host, rounds to float32, uploads to the
 * CURRENT device. Synchronous (cudaMalloc + cudaMemcpy); call outside hot loops.
 * Errors: modes[-1] > S[-1]/2+1 or modes[a] > S[a] (the reference's einsum
 * shape-errors there). 2*modes[a] > S[a] is accepted with the reference's
 * "later corner block overwrites" semantics. */
int pdeqx_spectral_plan_create(const pdeqx_spectral_desc* desc, pdeqx_spectral_plan** plan);
void pdeqx_spectral_plan_destroy(pdeqx_spectral_plan* plan);
/* bytes of scratch the fwd / bwd calls need (max over both) */
size_t pdeqx_spectral_workspace_bytes(const pdeqx_spectral_plan* plan);
/* number of floats of the saved tensors: x_hat = [B, c_in, modes...] complex,
 * z = [B, c_out, S_1..S_{D-1}, m_D] complex (input of the last inverse stage) */
size_t pdeqx_spectral_xhat_floats(const pdeqx_spectral_plan* plan);
size_t pdeqx_spectral_z_floats(const pdeqx_spectral_plan* plan);
/* which stages of this plan run on tcgen05 (bit 0: first-stage r2c DFT, bit 1: last-stage slab kernel,
 * bit 2: bypass weight gradient, bit 3: per-mode channel mix and its two adjoints -- from 16 samples per call and
 * 64 channels on; below that the CUDA-core kernel is the faster one); 0 in fp32 precision or when the shape does not
 * qualify */
int pdeqx_spectral_plan_tcgen05_stages(const pdeqx_spectral_plan* plan);

/* y[B,c_out,*S] = act( irfftn(mix(rfftn(x))) + (bypass_w . x + bypass_b) )
 *   w_re, w_im : [G, c_out, c_in, *modes] (G = 2^(ndim-1), corner order of
 *                generate_modes_slices, _spectral_conv.py:73-104)
 *   bypass_w   : [c_out, c_in] or NULL (plain SpectralConv); bypass_b [c_out] or NULL
 *   save_xhat, save_z : outputs kept for the backward pass, or NULL (inference)
 */
int pdeqx_spectral_block_fwd(const pdeqx_spectral_plan* plan, const float* x, const float* w_re,
                             const float* w_im, const float* bypass_w, const float* bypass_b,
                             int32_t activation, float* y, float* save_xhat, float* save_z,
                             void* workspace, size_t workspace_bytes, void* stream);

/* The same forward pass with the two pointwise ends of an architecture folded in (both optional; both need
 * pdeqx_spectral_block_rank1_ok(plan, 1, activation) == 1, ndim >= 2 and a bypass):
 *  - lift_u != NULL (then x must be NULL): the block's input is the pointwise 1 -> c_in lifting x[b,i,pixel] = lift_w[i] *
 *    lift_u[b,pixel] + lift_b[i] of a single-channel field (pdequinox/arch/_classic_fno.py:71-76 in front of the first
 *    ClassicSpectralBlock) and is never written: its first DFT stage is the lifting of the first stage of lift_u (the
 *    transform is linear) and the bypass term is rank one, (bypass_w . lift_w)[o] * lift_u + const[o], added in the last
 *    kernel's epilogue.
 *  - proj_out != NULL (then y may be NULL): the block's output only feeds the pointwise C -> 1 projection
 *    (pdequinox/arch/_classic_fno.py:78-83): proj_out[b,pixel] = sum_o proj_w[o] y[b,o,pixel] + proj_b[0] is reduced in
 *    the last kernel's epilogue and y is never written. */
int pdeqx_spectral_block_fwd_ex(const pdeqx_spectral_plan* plan, const float* x, const float* lift_u,
                                const float* lift_w, const float* lift_b, const float* proj_w, const float* proj_b,
                                float* proj_out, const float* w_re, const float* w_im, const float* bypass_w,
                                const float* bypass_b, int32_t activation, float* y, float* save_xhat, float* save_z,
                                void* workspace, size_t workspace_bytes, void* stream);

/* VJP of the call above (what jax.custom_vjp's bwd rule binds).
 *   gy [B,c_out,*S] cotangent of y; x, saved_xhat, saved_z from the forward.
 *   Outputs (overwritten, weight grads summed over the batch):
 *   gx [B,c_in,*S]; gw_re, gw_im [G,c_out,c_in,*modes]; g_bypass_w [c_out,c_in];
 *   g_bypass_b [c_out] (the last two may be NULL when bypass_w is NULL).
 *   gx may be NULL (first layer). scratch_full: [B,c_out,*S] floats (g_pre).
 */
int pdeqx_spectral_block_bwd(const pdeqx_spectral_plan* plan, const float* x, const float* gy,
                             const float* w_re, const float* w_im, const float* bypass_w,
                             const float* bypass_b, int32_t activation, const float* saved_xhat,
                             const float* saved_z, float* gx, float* gw_re, float* gw_im,
                             float* g_bypass_w, float* g_bypass_b, float* scratch_full,
                             void* workspace, size_t workspace_bytes, void* stream);

/* The same VJP with the two ends of an architecture folded in (both optional, both need
 * pdeqx_spectral_block_rank1_ok(plan, has_bypass, activation) == 1: tcgen05 slab kernel + fused activation):
 *  - gy_w != NULL: RANK-ONE cotangent gy_full[b,o,pixel] = gy_w[o] * gy[b,pixel] -- what reaches the last spectral block
 *    through the C -> 1 pointwise projection (pdequinox/arch/_classic_fno.py:78-83; the data gradient of
 *    pdequinox/conv/_pointwise_linear_conv.py:37-49 is exactly this outer product). gy is then [B,*S]; the full-size
 *    cotangent is never written or read, the kernel forms it in its epilogue.
 *  - lift_u != NULL (gx must be NULL): the block's input was produced from the single-channel field lift_u [B,*S] by the
 *    1 -> c_in pointwise lifting (pdequinox/arch/_classic_fno.py:71-76); instead of storing gx the last kernel
 *    accumulates that lifting's gradients lift_gw[i] = sum gx[b,i,pixel] lift_u[b,pixel], lift_gb[i] = sum gx (may be
 *    NULL). Outputs are overwritten. If additionally x == NULL the forward pass ran through
 *    pdeqx_spectral_block_fwd_ex with a lifted input: lift_w[i] * lift_u + lift_b[i] (lift_b may be NULL) was never
 *    materialised and the reverse pass works from lift_u, lift_w, lift_b alone.
 *  - proj_gw != NULL (needs gy_w): the forward pass projected the block's output without storing it
 *    (pdeqx_spectral_block_fwd_ex with proj_out); the projection's weight gradient proj_gw[o] = sum_{b,pixel}
 *    act(pre)[b,o,pixel] * gy[b,pixel] is accumulated from the activation recomputed in the cotangent pass. */
int pdeqx_spectral_block_rank1_ok(const pdeqx_spectral_plan* plan, int32_t has_bypass, int32_t activation);
int pdeqx_spectral_block_bwd_ex(const pdeqx_spectral_plan* plan, const float* x, const float* gy, const float* gy_w,
                                const float* lift_u, float* lift_gw, float* lift_gb, const float* lift_w,
                                const float* lift_b, float* proj_gw, const float* w_re,
                                const float* w_im, const float* bypass_w, const float* bypass_b,
                                int32_t activation, const float* saved_xhat, const float* saved_z, float* gx,
                                float* gw_re, float* gw_im, float* g_bypass_w, float* g_bypass_b,
                                float* scratch_full, void* workspace, size_t workspace_bytes, void* stream);

/* The VJP with every argument named (what the FFI handler fills from its operand list), plus CHAINING of two consecutive
 * ClassicSpectralBlocks of a Sequential (pdequinox/_sequential.py:111-116: block k-1's output is block k's input):
 *  - up_plan != NULL: the last kernel of this block's reverse pass does not store gx; in the same pass over the tiles it
 *    multiplies the input-gradient tile by act'(pre) of the UPSTREAM block (pre recomputed from up_saved_z, up_x and the
 *    upstream bypass, exactly as that block's own cotangent pass would) and stores up_gpre [B, C, *S], the cotangent of
 *    the upstream block's pre-activation. Needs pdeqx_spectral_block_chain_ok(plan, up_plan, up_activation) == 1 (tcgen05
 *    path, same batch / grid, all channel counts equal, 32 or 64) and gx == NULL. up_x == NULL: the upstream block's
 *    input was a folded 1 -> C lifting of up_lift_u (see pdeqx_spectral_block_fwd_ex).
 *  - gy_is_gpre != 0: `gy` already IS the cotangent of this block's pre-activation (the up_gpre of the chained call made
 *    for the block behind it): the activation-derivative pass is skipped; saved_z / scratch_full are not needed for it.
 * Every field not used must be zero (initialise the struct with {0}). */
typedef struct pdeqx_spectral_bwd_args {
  const float *x, *gy, *gy_w, *lift_u;
  float *lift_gw, *lift_gb;
  const float *lift_w, *lift_b;
  float* proj_gw;
  const float *w_re, *w_im, *bypass_w, *bypass_b;
  int32_t activation;
  const float *saved_xhat, *saved_z;
  float *gx, *gw_re, *gw_im, *g_bypass_w, *g_bypass_b, *scratch_full;
  void* workspace;
  size_t workspace_bytes;
  void* stream;
  int32_t gy_is_gpre;
  const pdeqx_spectral_plan* up_plan;
  const float *up_x, *up_saved_z, *up_bypass_w, *up_bypass_b;
  int32_t up_activation;
  const float *up_lift_u, *up_lift_w, *up_lift_b;
  float* up_gpre;
} pdeqx_spectral_bwd_args;
int pdeqx_spectral_block_bwd_args(const pdeqx_spectral_plan* plan, const pdeqx_spectral_bwd_args* args);
int pdeqx_spectral_block_chain_ok(const pdeqx_spectral_plan* plan, const pdeqx_spectral_plan* up_plan, int32_t up_activation);

/* ------------------------------------------------------------------------- *
 * PointwiseLinearConv (1x1 conv)
 *   replaces pdequinox/conv/_pointwise_linear_conv.py:6-53 (eqx.nn.Conv, k=1)
 * ------------------------------------------------------------------------- */
/* y[B,c_out,N] = act(w[c_out,c_in] . x[B,c_in,N] + b[c_out] + add[B,c_out,N]); b, add may be NULL */
int pdeqx_pointwise_fwd(int32_t batch, int32_t c_in, int32_t c_out, int64_t n_pixels, const float* x,
                        const float* w, const float* b, const float* add, int32_t activation,
                        float* y, void* stream);
/* gx = w^T . gy (gx may be NULL); gw[c_out,c_in] = sum_{b,n} gy x^T; gb[c_out] = sum gy (may be NULL).
 * Outputs are overwritten. */
int pdeqx_pointwise_bwd(int32_t batch, int32_t c_in, int32_t c_out, int64_t n_pixels, const float* x,
                        const float* w, const float* gy, float* gx, float* gw, float* gb,
                        void* stream);

/* ------------------------------------------------------------------------- *
 * PhysicsConv / MorePaddingConv
 *   replaces pdequinox/conv/_conv.py:169-219 with the padding of
 *   pdequinox/conv/_physics_conv.py:11-26; the halo is built in shared memory.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t ndim; /* 1..3 */
  int32_t batch;
  int32_t c_in, c_out, groups;
  int32_t spatial[PDEQX_MAX_DIMS]; /* input S */
  int32_t kernel[PDEQX_MAX_DIMS];
  int32_t stride[PDEQX_MAX_DIMS];
  int32_t dilation[PDEQX_MAX_DIMS];
  int32_t pad_lo[PDEQX_MAX_DIMS], pad_hi[PDEQX_MAX_DIMS];
  int32_t boundary; /* enum pdeqx_boundary */
  int32_t precision; /* enum pdeqx_precision: PDEQX_PREC_FP32 (CUDA-core fp32, every shape) or PDEQX_PREC_TF32 (tcgen05
                        implicit GEMM for the shapes that qualify: 2-D, k = 3, stride 1, W = 128, C = 64; fp32 otherwise) */
} pdeqx_conv_desc;

/* which passes of this convolution run on tcgen05 (bit 0: forward, bit 1: data gradient, bit 2: filter gradient);
 * 0 = CUDA-core fp32 path */
int pdeqx_conv_tcgen05_passes(const pdeqx_conv_desc* d);
/* bytes of caller-provided scratch the three passes below need (re-laid-out filter + per-CTA partial filter gradients
 * of the tcgen05 passes; 0 for the CUDA-core path, where `workspace` may be NULL). The calls keep nothing between
 * launches: one workspace per stream makes them re-entrant. */
size_t pdeqx_conv_workspace_bytes(const pdeqx_conv_desc* d);
/* how many forward / data-gradient launches of this process took the CTA-pair kernel (tcgen05 cta_group::2: two CTAs of one
 * TPC share one N = 192 MMA, each holding half of the filter): the C = 64 shapes whose row chains split into equal segments.
 * A diagnostic for tests and profiles; the one-CTA kernel covers the rest. */
int64_t pdeqx_conv_pair_launches(void);
/* output spatial size per axis: floor((s + lo + hi - d(k-1) - 1)/stride) + 1 */
int pdeqx_conv_out_spatial(const pdeqx_conv_desc* d, int32_t out_spatial[PDEQX_MAX_DIMS]);
/* y = act(corr(pad(x), w) + b + residual); w [c_out, c_in/groups, *k]; b [c_out]; b, residual may be NULL */
int pdeqx_conv_fwd(const pdeqx_conv_desc* d, const float* x, const float* w, const float* b,
                   const float* residual, int32_t activation, float* y, void* workspace, size_t workspace_bytes,
                   void* stream);
/* gx = ( adjoint-of-pad( corr^T(gy, w) ) + add ) * [relu_mask > 0], overwritten. add [B,c_in,*S] may be NULL (it carries
 * the cotangent of a residual connection into the same pass). relu_mask [B,c_in,*S] may be NULL: the OUTPUT of the relu
 * layer whose result this convolution read (pdequinox/blocks/_classic_res_block.py:112-121: conv_2(act(conv_1(x)))) --
 * its derivative is applied where the cotangent is produced instead of in an elementwise pass of its own. */
int pdeqx_conv_bwd_data(const pdeqx_conv_desc* d, const float* gy, const float* w, const float* add,
                        const float* relu_mask, float* gx, void* workspace, size_t workspace_bytes, void* stream);
/* gw[c_out,c_in/groups,*k] = sum_{b,p} gy . pad(x); gb[c_out] = sum gy (may be NULL); overwritten */
int pdeqx_conv_bwd_filter(const pdeqx_conv_desc* d, const float* x, const float* gy, float* gw,
                          float* gb, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * PhysicsConvTranspose / MorePaddingConvTranspose (ClassicUNet up-sampling)
 *   replaces pdequinox/conv/_conv.py:378-471 with the SAME padding of
 *   pdequinox/conv/_physics_conv.py:219-232. Reference-specific semantics: the
 *   weight [c_out, c_in/groups, *k] is used as is (no flip, no in/out swap) in
 *   a stride-1 correlation over the boundary-padded, stride-dilated input.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t ndim; /* 1..3 */
  int32_t batch;
  int32_t c_in, c_out, groups;
  int32_t spatial[PDEQX_MAX_DIMS]; /* input S */
  int32_t kernel[PDEQX_MAX_DIMS];
  int32_t stride[PDEQX_MAX_DIMS];  /* up-sampling factor (lhs dilation) */
  int32_t dilation[PDEQX_MAX_DIMS];
  int32_t pad_lo[PDEQX_MAX_DIMS], pad_hi[PDEQX_MAX_DIMS]; /* the module's `padding` */
  int32_t output_padding[PDEQX_MAX_DIMS];
  int32_t boundary; /* enum pdeqx_boundary */
} pdeqx_convT_desc;
int pdeqx_convT_out_spatial(const pdeqx_convT_desc* d, int32_t out_spatial[PDEQX_MAX_DIMS]);
/* y[B,c_out,*S_out] = corr(dilate(pad(x)), w) + b; b may be NULL */
int pdeqx_convT_fwd(const pdeqx_convT_desc* d, const float* x, const float* w, const float* b, float* y,
                    void* stream);
/* gx[B,c_in,*S] (overwritten) */
int pdeqx_convT_bwd_data(const pdeqx_convT_desc* d, const float* gy, const float* w, float* gx,
                         void* stream);
/* gw[c_out,c_in/groups,*k], gb[c_out] (may be NULL); overwritten */
int pdeqx_convT_bwd_filter(const pdeqx_convT_desc* d, const float* x, const float* gy, float* gw,
                           float* gb, void* stream);

/* Skip connection of Hierarchical.__call__ (pdequinox/_hierarchical.py:275):
 * out[B, c_a + c_b, n] = concatenate((a, b), channel axis), and its adjoint. */
int pdeqx_concat_channels(int32_t batch, int32_t c_a, int32_t c_b, int64_t n, const float* a,
                          const float* b, float* out, void* stream);
int pdeqx_split_channels(int32_t batch, int32_t c_a, int32_t c_b, int64_t n, const float* g, float* ga,
                         float* gb, void* stream);

/* ------------------------------------------------------------------------- *
 * GroupNorm (eqx.nn.GroupNorm, eps 1e-5, biased variance, per-channel affine)
 *   call sites: pdequinox/blocks/_dilated_res_block.py:88-99,143-146,
 *   _classic_double_conv_block.py:75-83, _classic_res_block.py:80-89
 * ------------------------------------------------------------------------- */
size_t pdeqx_groupnorm_workspace_bytes(int32_t batch, int32_t channels);
/* y[B,C,n] = (x - mean_g) * rstd_g * weight[c] + bias[c]; stats [B,groups,2] = (mean, rstd) is an output
 * kept for the backward pass; weight/bias may be NULL; workspace: pdeqx_groupnorm_workspace_bytes. */
int pdeqx_groupnorm_fwd(int32_t batch, int32_t channels, int32_t groups, int64_t n, const float* x,
                        const float* weight, const float* bias, float eps, float* y, float* stats,
                        void* workspace, void* stream);
/* gx, gweight[C], gbias[C] (overwritten; the last two may be NULL); scratch: 2*B*C floats. relu_mask [B,C,n] may be NULL:
 * gx is zeroed where relu_mask <= 0 -- the relu' of the layer whose OUTPUT this norm read (norm -> conv -> act chains,
 * pdequinox/blocks/_dilated_res_block.py:141-151), applied here instead of in an elementwise pass of its own. */
int pdeqx_groupnorm_bwd(int32_t batch, int32_t channels, int32_t groups, int64_t n, const float* x,
                        const float* weight, const float* stats, const float* gy, const float* relu_mask, float* gx,
                        float* gweight, float* gbias, float* scratch, void* stream);

/* ------------------------------------------------------------------------- *
 * GroupNorm folded into the convolutions around it ("stats in the producer's epilogue, normalise on load")
 *   replaces the norm -> conv -> act chain of pdequinox/blocks/_dilated_res_block.py:141-151 (eqx.nn.GroupNorm with
 *   groups = 1 in front of a periodic PhysicsConv) without the normalised tensor ever being written or read:
 *     forward   conv(gamma (h - mu_b) rstd_b + beta) = rstd_b conv(h; W gamma) + const[b, o]: gamma rides in the prepared
 *               filter, rstd_b and the constant in the epilogue, which also sums y, y^2 per sample for the NEXT norm;
 *     filter    gradient against the virtual normalised input: the cotangent operand is scaled by rstd_b on its way into
 *               tensor memory, gamma / beta / mu enter as a rank-one correction when the per-CTA partial sums are reduced;
 *     data      v = corr^T(gy, w) with sum gamma v and sum gamma v h per sample from the epilogue, then ONE elementwise
 *               pass (pdeqx_groupnorm_bwd_fused) forms the norm's input gradient, relu' of the layer in front and the
 *               norm's parameter gradients.
 *   Accepted shapes: pdeqx_conv_gn_fusable(d) != 0 (the tcgen05 CTA-pair shapes -- 2-D, k = 3, stride 1, W = 128,
 *   C = 64, H % dilation == 0, TF32 precision -- with periodic boundary: only there is the norm's shift a per-sample
 *   constant). Everything else takes pdeqx_groupnorm_fwd/bwd + pdeqx_conv_*.
 *   parts: [B][slots][2] fp32 partial sums, slots = pdeqx_conv_gn_slots(d) for the buffers the convolution passes write.
 * ------------------------------------------------------------------------- */
int pdeqx_conv_gn_fusable(const pdeqx_conv_desc* d);
int pdeqx_conv_gn_slots(const pdeqx_conv_desc* d);
/* y = act(corr(pad(gn(x)), w) + b), gn = GroupNorm(groups = 1) with statistics gn_stats [B][2] = (mean, rstd) of x and
 * affine gn_weight / gn_bias [c_in] (NULL: 1 / 0). out_parts (may be NULL): per-sample partial sums (y, y^2). */
int pdeqx_conv_gn_fwd(const pdeqx_conv_desc* d, const float* x, const float* w, const float* b, const float* gn_stats,
                      const float* gn_weight, const float* gn_bias, int32_t activation, float* y, float* out_parts,
                      void* workspace, size_t workspace_bytes, void* stream);
/* gw = sum gy . pad(gn(x)), gb = sum gy (may be NULL); x is the norm's INPUT */
int pdeqx_conv_gn_bwd_filter(const pdeqx_conv_desc* d, const float* x, const float* gy, const float* gn_stats,
                             const float* gn_weight, const float* gn_bias, float* gw, float* gb, void* workspace,
                             size_t workspace_bytes, void* stream);
/* v = adjoint-of-pad(corr^T(gy, w)) (the cotangent of gn(h)); parts [B][slots][2] = per-sample partial sums
 * (gamma_c v, gamma_c v h) */
int pdeqx_conv_gn_bwd_data(const pdeqx_conv_desc* d, const float* gy, const float* w, const float* gn_weight, const float* h,
                           float* v, float* parts, void* workspace, size_t workspace_bytes, void* stream);
/* statistics only: stats [B,groups,2] = (mean, rstd); workspace: pdeqx_groupnorm_workspace_bytes */
int pdeqx_groupnorm_stats(int32_t batch, int32_t channels, int32_t groups, int64_t n, const float* x, float eps,
                          float* stats, void* workspace, void* stream);
/* groups = 1: stats [B][2] from per-sample partial sums (x, x^2) over `count` elements per sample */
int pdeqx_groupnorm_stats_from_parts(int32_t batch, int32_t slots, int64_t count, const float* parts, float eps,
                                     float* stats, void* stream);
/* out = a + b (the residual join of a block, [B][per_sample]) + per-sample partial sums (out, out^2) in parts [B][slots][2];
 * per_sample % (4 slots) == 0 */
int pdeqx_add_parts(int32_t batch, int64_t per_sample, int32_t slots, const float* a, const float* b, float* out,
                    float* parts, void* stream);
/* gx = rstd_b (gamma_c v - s1_b - xhat s2_b) [* (h > 0) if relu_mask_h] [+ add], gweight, gbias (may be NULL; overwritten);
 * parts / slots as written by pdeqx_conv_gn_bwd_data; scratch: 2*B*C floats; n % 4 == 0 */
int pdeqx_groupnorm_bwd_fused(int32_t batch, int32_t channels, int64_t n, const float* h, const float* v,
                              const float* weight, const float* stats, const float* parts, int32_t slots,
                              int32_t relu_mask_h, const float* add, float* gx, float* gweight, float* gbias,
                              float* scratch, void* stream);

/* ------------------------------------------------------------------------- *
 * Elementwise pieces of the block / training step
 * ------------------------------------------------------------------------- */
/* gx = gy * act'(pre) (gelu'/relu' evaluated on the PRE-activation) */
int pdeqx_activation_bwd(int64_t n, const float* pre, const float* gy, int32_t activation, float* gx,
                         void* stream);
/* y = act(x) */
int pdeqx_activation_fwd(int64_t n, const float* x, int32_t activation, float* y, void* stream);
/* out = a + b (residual joins that cannot be fused into an epilogue); out may alias a or b */
int pdeqx_add(int64_t n, const float* a, const float* b, float* out, void* stream);
/* loss[0] = mean((pred-target)^2) (jnp.mean over all elements, README.md:74-76);
 * gpred = grad_scale * 2/n * (pred-target) (may be NULL). partials: >= 1024 floats scratch. */
int pdeqx_mse_loss(int64_t n, const float* pred, const float* target, float grad_scale, float* loss,
                   float* gpred, float* partials, void* stream);
/* optax.adam update on a flat arena: g is first multiplied by grad_scale (1/world_size).
 * step is the 1-based update count (bias correction). */
int pdeqx_adam_step(int64_t n, float* param, const float* grad, float* mu, float* nu, float lr,
                    float b1, float b2, float eps, int32_t step, float grad_scale, void* stream);
/* Same update with the step count kept ON THE DEVICE, so that a captured CUDA graph of the whole train
 * step can be replayed: state[0] holds the update count (int32 bit pattern, start at 0) and is
 * incremented by the call; state[1..2] receive the bias corrections. state: 4 floats. */
int pdeqx_adam_step_dev(int64_t n, float* param, const float* grad, float* mu, float* nu, float lr,
                        float b1, float b2, float eps, float* state, float grad_scale, void* stream);

/* ------------------------------------------------------------------------- *
 * Data-parallel exchange over NVLink peer memory (one process per GPU, one node):
 * reduce-scatter of the gradient arenas + Adam on the rank's 1/world shard + all-gather of the new parameters in ONE
 * kernel, instead of an all-reduce followed by a replicated optimiser pass. The reference has no multi-device code
 * (SURVEY.md 2.2); this is the exchange step of SURVEY.md 8(e).
 *   - pdeqx_dp_alloc: cudaMalloc'ed (zeroed) buffer of its own -- exportable; hold the parameter arena, the gradient arena
 *     and the signal pad (pdeqx_dp_pad_bytes) of every rank in such buffers;
 *   - pdeqx_dp_export / _import: CUDA IPC handle of such a buffer / a pointer to a peer's buffer valid on this device
 *     (the 64-byte handles travel through whatever the host uses for rendezvous, e.g. torch.distributed.all_gather_object);
 *     any other way of obtaining peer-mapped (and multicast-mapped) pointers works as well, e.g. torch's symmetric memory;
 *   - pdeqx_dp_adam_step: grads[r], params[r], pads[r] for r < world are every rank's buffers as seen from THIS device
 *     (the rank's own ones included). mu_shard / nu_shard: pdeqx_dp_shard_floats(n, world) floats each (Adam moments of the
 *     rank's shard only). state: 4 floats as pdeqx_adam_step_dev. loss_io (may be NULL): in = this rank's loss, out = mean
 *     over the ranks. Enqueues only; ranks synchronise through flags in the signal pads, a wait longer than ~3 s sets an
 *     error word that pdeqx_dp_check reports (it synchronises the stream).
 * ------------------------------------------------------------------------- */
#define PDEQX_MAX_RANKS 8
typedef struct {
  int32_t world, rank;
  float* grads[PDEQX_MAX_RANKS];
  float* params[PDEQX_MAX_RANKS];
  uint32_t* pads[PDEQX_MAX_RANKS];
  /* optional NVSwitch multicast (NVLS) addresses of the SAME gradient / parameter arenas (one multicast object bound to
   * every rank's buffer; NULL = not available): the kernel then sums the N gradient slices with ONE multimem.ld_reduce
   * per 16 bytes (the switch reduces: 1/N of the arena crosses this GPU's links instead of (N-1)/N) and publishes the
   * new parameters with ONE multimem.st (the switch replicates). */
  float* mc_grads;
  float* mc_params;
} pdeqx_dp_peers;
int pdeqx_dp_alloc(size_t bytes, void** out);
int pdeqx_dp_free(void* p);
int pdeqx_dp_export(const void* p, uint8_t handle[64]);
int pdeqx_dp_import(const uint8_t handle[64], void** out);
int pdeqx_dp_unimport(void* p);
size_t pdeqx_dp_pad_bytes(void);
int64_t pdeqx_dp_shard_floats(int64_t n, int32_t world);
int pdeqx_dp_adam_step(const pdeqx_dp_peers* peers, int64_t n, float* mu_shard, float* nu_shard, float lr, float b1,
                       float b2, float eps, float* state, float* loss_io, void* stream);
int pdeqx_dp_check(const pdeqx_dp_peers* peers, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDEQX_B200_H_ */
