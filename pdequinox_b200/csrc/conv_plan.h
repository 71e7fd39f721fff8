// Internal: resolved geometry of one PhysicsConv call, shared by conv.cu (any-shape fp32 path)
// and tc_conv.cu (tcgen05 implicit GEMM).
#pragma once
#include "common.cuh"

namespace pdeqx {

struct ConvGeom {
  int ndim, batch, c_in, c_out, groups, boundary, taps, max_pre, precision;
  int spatial[PDEQX_MAX_DIMS], out[PDEQX_MAX_DIMS];
  int kernel[PDEQX_MAX_DIMS], stride[PDEQX_MAX_DIMS], dilation[PDEQX_MAX_DIMS];
  int pad_lo[PDEQX_MAX_DIMS], pad_hi[PDEQX_MAX_DIMS];
  int in_stride[PDEQX_MAX_DIMS], out_stride[PDEQX_MAX_DIMS];  // element strides inside one channel
  int64_t n_in, n_out;                                        // points per channel
};

int conv_geometry(const pdeqx_conv_desc* d, ConvGeom* out);

// gx += the part of the data gradient that arrives through wrapped / reflected halo positions (conv.cu, CUDA cores)
int conv_bwd_data_mirror_part(const ConvGeom& g, const float* gy, const float* w, float* gx, cudaStream_t s);

// tcgen05 implicit-GEMM path (tc_conv.cu)
bool tc_conv_available(const ConvGeom& g);
bool tc_conv_bwd_data_available(const ConvGeom& g);
bool tc_conv_bwd_filter_available(const ConvGeom& g);
// scratch of the tcgen05 passes (0 when the geometry does not take them): re-laid-out filter + partial filter gradients
size_t tc_conv_workspace_bytes(const ConvGeom& g);
int tc_conv_fwd(const ConvGeom& g, const float* x, const float* w, const float* b, const float* residual, int act,
                float* y, void* ws, size_t ws_bytes, cudaStream_t s);
// relu_mask: applied in the kernel's epilogue except for the neumann boundary (two-pass data gradient: the caller masks)
int tc_conv_bwd_data(const ConvGeom& g, const float* gy, const float* w, const float* add, const float* relu_mask, float* gx,
                     void* ws, size_t ws_bytes, cudaStream_t s);
int tc_conv_bwd_filter(const ConvGeom& g, const float* x, const float* gy, float* gw, float* gb, void* ws, size_t ws_bytes,
                       cudaStream_t s);

// GroupNorm (groups = 1) folded into the three passes of a periodic tcgen05 convolution (tc_conv.cu)
bool tc_conv_gn_fusable(const ConvGeom& g);
int tc_conv_gn_slots(const ConvGeom& g);
int tc_conv_gn_fwd(const ConvGeom& g, const float* x, const float* w, const float* b, const float* stats, const float* gamma,
                   const float* beta, int act, float* y, float* out_parts, void* ws, size_t ws_bytes, cudaStream_t s);
int tc_conv_gn_bwd_filter(const ConvGeom& g, const float* x, const float* gy, const float* stats, const float* gamma,
                          const float* beta, float* gw, float* gb, void* ws, size_t ws_bytes, cudaStream_t s);
int tc_conv_gn_bwd_data(const ConvGeom& g, const float* gy, const float* w, const float* gamma, const float* h, float* v,
                        float* parts, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace pdeqx
