// XLA FFI custom-call handlers over the C ABI of libpdeqx_b200.so (include/pdeqx_b200.h).
//
// This is the binding BASELINE.json's north star asks for: "Host code stays in Python/JAX and calls the CUDA kernels
// through a thin C-ABI layer registered as XLA FFI custom calls via jax.ffi, with custom_vjp for the backward pass".
// Each handler replaces one reference call site (single sample under jax.vmap -> the batch axis is the kernels' own):
//
//   pdeqx_ffi_spectral_block_fwd / _bwd : SpectralConv.__call__ (pdequinox/conv/_spectral_conv.py:65-66,107-136) and, with a
//                                         bypass operand, ClassicSpectralBlock.__call__ (blocks/_classic_spectral_block.py:72-75)
//   pdeqx_ffi_conv_fwd / _bwd           : MorePaddingConv.__call__ (pdequinox/conv/_conv.py:169-219) + PhysicsConv's SAME padding
//   pdeqx_ffi_convT_fwd / _bwd          : MorePaddingConvTranspose.__call__ (pdequinox/conv/_conv.py:378-471)
//   pdeqx_ffi_pointwise_fwd / _bwd      : PointwiseLinearConv (pdequinox/conv/_pointwise_linear_conv.py:37-49)
//   pdeqx_ffi_groupnorm_fwd / _bwd      : eqx.nn.GroupNorm call sites (blocks/_dilated_res_block.py:88-99)
//   pdeqx_ffi_adam                      : optax.adam(3e-4) + eqx.apply_updates (README.md:78-86), in place (aliased operands)
//
// Only the plain-C call-frame ABI (xla/ffi/api/c_api.h) is used -- no C++ template layer -- so the shim has no build
// dependency beyond that one header. Build: `make xla_ffi` (Makefile): against jaxlib's header when `import jax.ffi`
// works, else against the minimal restatement under stub/ (compile check + the call-frame tests of this repository).
// The JAX side (registration, custom_vjp, custom_vmap) is pdequinox_b200/jax_ffi/.
//
// Contract of every handler: XLA owns all buffers; a handler only ENQUEUES on the stream XLA hands it, never
// synchronises, never allocates device memory at execute time (scratch arrives as an extra result buffer; spectral
// plans -- twiddle tables -- are created in the INITIALIZE stage, or lazily on the first execution outside stream
// capture) and never throws. Absent optional operands are passed as zero-element buffers.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "xla/ffi/api/c_api.h"

#include "pdeqx_b200.h"

extern "C" int pdeqx_current_device(void);  // libpdeqx_b200.so (core.cu): cudaGetDevice of the calling thread, -1 on error

namespace {

// ------------------------------------------------------------------------------------------
// call-frame access
// ------------------------------------------------------------------------------------------
XLA_FFI_Error* make_error(const XLA_FFI_Api* api, XLA_FFI_Error_Code code, const char* msg) {
  XLA_FFI_Error_Create_Args a;
  memset(&a, 0, sizeof(a));
  a.struct_size = sizeof(a);
  a.message = msg;
  a.errc = code;
  return api->XLA_FFI_Error_Create(&a);
}

struct Frame {
  XLA_FFI_CallFrame* f;
  char msg[640];
  explicit Frame(XLA_FFI_CallFrame* frame) : f(frame) { msg[0] = 0; }

  XLA_FFI_Error* fail(XLA_FFI_Error_Code code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
    snprintf(msg, sizeof(msg), fmt, a, b, c);
    return make_error(f->api, code, msg);
  }
  // status of a libpdeqx_b200 call -> XLA error (the library's thread-local message travels with it)
  XLA_FFI_Error* status(int rc, const char* what) {
    if (rc == PDEQX_OK) return nullptr;
    const XLA_FFI_Error_Code code = rc == PDEQX_ERR_INVALID_ARGUMENT ? XLA_FFI_Error_Code_INVALID_ARGUMENT
                                    : rc == PDEQX_ERR_UNSUPPORTED    ? XLA_FFI_Error_Code_UNIMPLEMENTED
                                    : rc == PDEQX_ERR_WORKSPACE      ? XLA_FFI_Error_Code_RESOURCE_EXHAUSTED
                                                                     : XLA_FFI_Error_Code_INTERNAL;
    snprintf(msg, sizeof(msg), "%s: %s", what, pdeqx_last_error());
    return make_error(f->api, code, msg);
  }

  bool check_counts(int n_args, int n_rets, XLA_FFI_Error** err, const char* name) {
    if (f->args.size != n_args || f->rets.size != n_rets) {
      *err = fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: expected %lld operands / %lld results", name, n_args, n_rets);
      return false;
    }
    for (int64_t i = 0; i < f->args.size; ++i)
      if (f->args.types[i] != XLA_FFI_ArgType_BUFFER || ((XLA_FFI_Buffer*)f->args.args[i])->dtype != XLA_FFI_DataType_F32) {
        *err = fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: operand %lld is not a float32 buffer", name, i);
        return false;
      }
    for (int64_t i = 0; i < f->rets.size; ++i)
      if (f->rets.types[i] != XLA_FFI_RetType_BUFFER || ((XLA_FFI_Buffer*)f->rets.rets[i])->dtype != XLA_FFI_DataType_F32) {
        *err = fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: result %lld is not a float32 buffer", name, i);
        return false;
      }
    return true;
  }
  const XLA_FFI_Buffer* arg(int i) const { return (const XLA_FFI_Buffer*)f->args.args[i]; }
  const XLA_FFI_Buffer* ret(int i) const { return (const XLA_FFI_Buffer*)f->rets.rets[i]; }
  static int64_t count(const XLA_FFI_Buffer* b) {
    int64_t n = 1;
    for (int64_t a = 0; a < b->rank; ++a) n *= b->dims[a];
    return n;
  }
  // optional operand: a zero-element buffer stands for "absent"
  const float* opt(int i) const { return count(arg(i)) == 0 ? nullptr : (const float*)arg(i)->data; }
  const float* in(int i) const { return (const float*)arg(i)->data; }
  float* out(int i) const { return (float*)ret(i)->data; }
  float* opt_out(int i) const { return count(ret(i)) == 0 ? nullptr : (float*)ret(i)->data; }
  size_t out_bytes(int i) const { return (size_t)count(ret(i)) * sizeof(float); }

  int find_attr(const char* name) const {
    const size_t n = strlen(name);
    for (int64_t i = 0; i < f->attrs.size; ++i)
      if (f->attrs.names[i]->len == n && memcmp(f->attrs.names[i]->ptr, name, n) == 0) return (int)i;
    return -1;
  }
  static bool scalar_as_double(const XLA_FFI_Scalar* s, double* v) {
    switch (s->dtype) {
      case XLA_FFI_DataType_S32: *v = *(const int32_t*)s->value; return true;
      case XLA_FFI_DataType_S64: *v = (double)*(const int64_t*)s->value; return true;
      case XLA_FFI_DataType_U32: *v = *(const uint32_t*)s->value; return true;
      case XLA_FFI_DataType_U64: *v = (double)*(const uint64_t*)s->value; return true;
      case XLA_FFI_DataType_F32: *v = *(const float*)s->value; return true;
      case XLA_FFI_DataType_F64: *v = *(const double*)s->value; return true;
      case XLA_FFI_DataType_PRED: *v = *(const uint8_t*)s->value; return true;
      default: return false;
    }
  }
  // scalar attribute (any integer / float width jax.ffi may have chosen for a Python or NumPy scalar)
  bool attr_number(const char* name, double* v) const {
    const int i = find_attr(name);
    if (i < 0 || f->attrs.types[i] != XLA_FFI_AttrType_SCALAR) return false;
    return scalar_as_double((const XLA_FFI_Scalar*)f->attrs.attributes[i], v);
  }
  // integer array attribute (np.int32 / np.int64 array) into out[0..PDEQX_MAX_DIMS); returns the length or -1
  int attr_ints(const char* name, int32_t out[PDEQX_MAX_DIMS], int32_t fill) const {
    for (int a = 0; a < PDEQX_MAX_DIMS; ++a) out[a] = fill;
    const int i = find_attr(name);
    if (i < 0 || f->attrs.types[i] != XLA_FFI_AttrType_ARRAY) return -1;
    const XLA_FFI_Array* arr = (const XLA_FFI_Array*)f->attrs.attributes[i];
    if (arr->size > PDEQX_MAX_DIMS) return -1;
    for (size_t a = 0; a < arr->size; ++a) {
      if (arr->dtype == XLA_FFI_DataType_S64) out[a] = (int32_t)((const int64_t*)arr->data)[a];
      else if (arr->dtype == XLA_FFI_DataType_S32) out[a] = ((const int32_t*)arr->data)[a];
      else return -1;
    }
    return (int)arr->size;
  }
  bool stream(void** s, XLA_FFI_Error** err) {
    XLA_FFI_Stream_Get_Args a;
    memset(&a, 0, sizeof(a));
    a.struct_size = sizeof(a);
    a.ctx = f->ctx;
    *err = f->api->XLA_FFI_Stream_Get(&a);
    *s = a.stream;
    return *err == nullptr;
  }
};

// A handler called with the metadata extension only reports its API version and traits (how XLA validates a handler at
// registration time). Every handler below only enqueues on the given stream -> command-buffer (CUDA graph) compatible.
bool answer_metadata(XLA_FFI_CallFrame* f) {
  if (f->extension_start == nullptr || f->extension_start->type != XLA_FFI_Extension_Metadata) return false;
  XLA_FFI_Metadata_Extension* ext = (XLA_FFI_Metadata_Extension*)f->extension_start;
  ext->metadata->api_version.major_version = XLA_FFI_API_MAJOR;
  ext->metadata->api_version.minor_version = XLA_FFI_API_MINOR;
  ext->metadata->traits = XLA_FFI_HANDLER_TRAITS_COMMAND_BUFFER_COMPATIBLE;
  return true;
}

#define PDEQX_FFI_PROLOGUE(NAME, N_ARGS, N_RETS)                                  \
  if (answer_metadata(call_frame)) return nullptr;                                \
  Frame fr(call_frame);                                                           \
  XLA_FFI_Error* err = nullptr;                                                   \
  if (!fr.check_counts(N_ARGS, N_RETS, &err, NAME)) return err;                   \
  const bool execute = call_frame->stage == XLA_FFI_ExecutionStage_EXECUTE;       \
  void* stream = nullptr;                                                         \
  if (execute && !fr.stream(&stream, &err)) return err

#define PDEQX_FFI_NUMBER(VAR, NAME, OP)                                                                              \
  double VAR = 0;                                                                                                     \
  if (!fr.attr_number(NAME, &VAR)) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: scalar attribute '" NAME "' is missing", OP)

// ------------------------------------------------------------------------------------------
// spectral plans: twiddle tables per (device, descriptor), created once, immutable afterwards
// ------------------------------------------------------------------------------------------
std::mutex g_plan_mu;
std::map<std::string, pdeqx_spectral_plan*> g_plans;

int plan_for(const pdeqx_spectral_desc& d, const pdeqx_spectral_plan** out) {
  const int dev = pdeqx_current_device();
  std::string key((const char*)&d, sizeof(d));
  key.append((const char*)&dev, sizeof(dev));
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    pdeqx_spectral_plan* p = nullptr;
    const int rc = pdeqx_spectral_plan_create(&d, &p);  // cudaMalloc + cudaMemcpy: INITIALIZE stage, outside stream capture
    if (rc != PDEQX_OK) return rc;
    it = g_plans.emplace(key, p).first;
  }
  *out = it->second;
  return PDEQX_OK;
}

// x [B, C_in, *S], weights [G, C_out, C_in, *modes] (call-time interpretation, _spectral_conv.py:129-134)
bool spectral_desc(const XLA_FFI_Buffer* x, const XLA_FFI_Buffer* w, int precision, pdeqx_spectral_desc* d) {
  memset(d, 0, sizeof(*d));
  const int D = (int)x->rank - 2;
  if (D < 1 || D > PDEQX_MAX_DIMS || w->rank != D + 3) return false;
  d->ndim = D;
  d->batch = (int32_t)x->dims[0];
  d->c_in = (int32_t)x->dims[1];
  d->c_out = (int32_t)w->dims[1];
  if (w->dims[2] != d->c_in || w->dims[0] != (1 << (D - 1))) return false;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    d->spatial[a] = a < D ? (int32_t)x->dims[2 + a] : 1;
    d->modes[a] = a < D ? (int32_t)w->dims[3 + a] : 1;
  }
  d->precision = precision;
  return true;
}

// x [B, C_in, *S], w [C_out, C_in / groups, *k] + the static attributes of the module
bool conv_desc(Frame& fr, const XLA_FFI_Buffer* x, const XLA_FFI_Buffer* w, pdeqx_conv_desc* d) {
  memset(d, 0, sizeof(*d));
  const int D = (int)x->rank - 2;
  if (D < 1 || D > PDEQX_MAX_DIMS || w->rank != D + 2) return false;
  double groups, boundary, precision;
  if (!fr.attr_number("groups", &groups) || !fr.attr_number("boundary", &boundary) || !fr.attr_number("precision", &precision))
    return false;
  d->ndim = D;
  d->batch = (int32_t)x->dims[0];
  d->c_in = (int32_t)x->dims[1];
  d->c_out = (int32_t)w->dims[0];
  d->groups = (int32_t)groups;
  d->boundary = (int32_t)boundary;
  d->precision = (int32_t)precision;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    d->spatial[a] = a < D ? (int32_t)x->dims[2 + a] : 1;
    d->kernel[a] = a < D ? (int32_t)w->dims[2 + a] : 1;
  }
  return fr.attr_ints("stride", d->stride, 1) == D && fr.attr_ints("dilation", d->dilation, 1) == D &&
         fr.attr_ints("pad_lo", d->pad_lo, 0) == D && fr.attr_ints("pad_hi", d->pad_hi, 0) == D;
}

bool convT_desc(Frame& fr, const XLA_FFI_Buffer* x, const XLA_FFI_Buffer* w, pdeqx_convT_desc* d) {
  memset(d, 0, sizeof(*d));
  const int D = (int)x->rank - 2;
  if (D < 1 || D > PDEQX_MAX_DIMS || w->rank != D + 2) return false;
  double groups, boundary;
  if (!fr.attr_number("groups", &groups) || !fr.attr_number("boundary", &boundary)) return false;
  d->ndim = D;
  d->batch = (int32_t)x->dims[0];
  d->c_in = (int32_t)x->dims[1];
  d->c_out = (int32_t)w->dims[0];
  d->groups = (int32_t)groups;
  d->boundary = (int32_t)boundary;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    d->spatial[a] = a < D ? (int32_t)x->dims[2 + a] : 1;
    d->kernel[a] = a < D ? (int32_t)w->dims[2 + a] : 1;
  }
  return fr.attr_ints("stride", d->stride, 1) == D && fr.attr_ints("dilation", d->dilation, 1) == D &&
         fr.attr_ints("pad_lo", d->pad_lo, 0) == D && fr.attr_ints("pad_hi", d->pad_hi, 0) == D &&
         fr.attr_ints("output_padding", d->output_padding, 0) == D;
}

int64_t pixels(const XLA_FFI_Buffer* x) {  // product of the spatial dims of [B, C, *S]
  int64_t n = 1;
  for (int64_t a = 2; a < x->rank; ++a) n *= x->dims[a];
  return n;
}

}  // namespace

// ==========================================================================================
// trace-time helper (called through ctypes by pdequinox_b200/jax_ffi/_ffi.py to size the result buffers of the spectral ops):
// sizes[0..2] = floats of x_hat, floats of z, BYTES of the workspace. Creates (and caches) the plan on the current device.
// ==========================================================================================
extern "C" int pdeqx_ffi_spectral_sizes(const pdeqx_spectral_desc* d, size_t sizes[3]) {
  const pdeqx_spectral_plan* plan = nullptr;
  const int rc = plan_for(*d, &plan);
  if (rc != PDEQX_OK) return rc;
  sizes[0] = pdeqx_spectral_xhat_floats(plan);
  sizes[1] = pdeqx_spectral_z_floats(plan);
  sizes[2] = pdeqx_spectral_workspace_bytes(plan);
  return PDEQX_OK;
}

// ==========================================================================================
// SpectralConv / ClassicSpectralBlock
//   operands: x [B,Ci,*S], w_re, w_im [G,Co,Ci,*M], bypass_w [Co,Ci,1..] or (0,), bypass_b [Co,1..] or (0,)
//   attrs   : activation (pdeqx_activation), precision (pdeqx_precision)
//   results : y [B,Co,*S], x_hat (flat, saved), z (flat, saved), workspace (flat scratch)
// register with {"initialize": this, "execute": this}: the plan (twiddle tables) is built in the INITIALIZE stage
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_spectral_block_fwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_spectral_block_fwd", 5, 4);
  PDEQX_FFI_NUMBER(activation, "activation", "pdeqx_spectral_block_fwd");
  PDEQX_FFI_NUMBER(precision, "precision", "pdeqx_spectral_block_fwd");
  pdeqx_spectral_desc d;
  if (!spectral_desc(fr.arg(0), fr.arg(1), (int)precision, &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: x must be [B, C_in, *S] and the weights [2^(D-1), C_out, C_in, *modes]",
                   "pdeqx_spectral_block_fwd");
  const pdeqx_spectral_plan* plan = nullptr;
  if ((err = fr.status(plan_for(d, &plan), "pdeqx_spectral_plan_create"))) return err;
  if (!execute) return nullptr;
  return fr.status(pdeqx_spectral_block_fwd(plan, fr.in(0), fr.in(1), fr.in(2), fr.opt(3), fr.opt(4), (int32_t)activation, fr.out(0),
                                            fr.out(1), fr.out(2), fr.out(3), fr.out_bytes(3), stream),
                   "pdeqx_spectral_block_fwd");
}

//   operands: x, gy [B,Co,*S], w_re, w_im, bypass_w or (0,), bypass_b or (0,), x_hat, z (saved by the forward call)
//   results : gx [B,Ci,*S], gw_re, gw_im (batch-reduced), g_bypass_w or (0,), g_bypass_b or (0,), scratch [B,Co,*S], workspace
extern "C" XLA_FFI_Error* pdeqx_ffi_spectral_block_bwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_spectral_block_bwd", 8, 7);
  PDEQX_FFI_NUMBER(activation, "activation", "pdeqx_spectral_block_bwd");
  PDEQX_FFI_NUMBER(precision, "precision", "pdeqx_spectral_block_bwd");
  pdeqx_spectral_desc d;
  if (!spectral_desc(fr.arg(0), fr.arg(2), (int)precision, &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: x must be [B, C_in, *S] and the weights [2^(D-1), C_out, C_in, *modes]",
                   "pdeqx_spectral_block_bwd");
  const pdeqx_spectral_plan* plan = nullptr;
  if ((err = fr.status(plan_for(d, &plan), "pdeqx_spectral_plan_create"))) return err;
  if (!execute) return nullptr;
  pdeqx_spectral_bwd_args a;
  memset(&a, 0, sizeof(a));
  a.x = fr.in(0);
  a.gy = fr.in(1);
  a.w_re = fr.in(2);
  a.w_im = fr.in(3);
  a.bypass_w = fr.opt(4);
  a.bypass_b = fr.opt(5);
  a.activation = (int32_t)activation;
  a.saved_xhat = fr.in(6);
  a.saved_z = fr.in(7);
  a.gx = fr.out(0);
  a.gw_re = fr.out(1);
  a.gw_im = fr.out(2);
  a.g_bypass_w = fr.opt_out(3);
  a.g_bypass_b = fr.opt_out(4);
  a.scratch_full = fr.out(5);
  a.workspace = fr.out(6);
  a.workspace_bytes = fr.out_bytes(6);
  a.stream = stream;
  return fr.status(pdeqx_spectral_block_bwd_args(plan, &a), "pdeqx_spectral_block_bwd");
}

// ==========================================================================================
// PhysicsConv / MorePaddingConv
//   fwd operands: x [B,Ci,*S], w [Co,Ci/g,*k], b [Co,1..] or (0,), residual [B,Co,*S_out] or (0,)
//       attrs   : stride, dilation, pad_lo, pad_hi (int arrays of length D), boundary, groups, activation, precision
//       results : y [B,Co,*S_out], workspace (pdeqx_conv_workspace_bytes; may be one element when it is 0)
//   bwd operands: x, gy [B,Co,*S_out], w, add [B,Ci,*S] or (0,)       results: gx, gw, gb or (0,), workspace
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_conv_fwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_conv_fwd", 4, 2);
  PDEQX_FFI_NUMBER(activation, "activation", "pdeqx_conv_fwd");
  pdeqx_conv_desc d;
  if (!conv_desc(fr, fr.arg(0), fr.arg(1), &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: bad operand ranks or missing static attributes", "pdeqx_conv_fwd");
  if (!execute) return nullptr;
  return fr.status(pdeqx_conv_fwd(&d, fr.in(0), fr.in(1), fr.opt(2), fr.opt(3), (int32_t)activation, fr.out(0), fr.out(1),
                                  fr.out_bytes(1), stream),
                   "pdeqx_conv_fwd");
}

extern "C" XLA_FFI_Error* pdeqx_ffi_conv_bwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_conv_bwd", 4, 4);
  pdeqx_conv_desc d;
  if (!conv_desc(fr, fr.arg(0), fr.arg(2), &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: bad operand ranks or missing static attributes", "pdeqx_conv_bwd");
  if (!execute) return nullptr;
  if ((err = fr.status(pdeqx_conv_bwd_filter(&d, fr.in(0), fr.in(1), fr.out(1), fr.opt_out(2), fr.out(3), fr.out_bytes(3), stream),
                       "pdeqx_conv_bwd_filter")))
    return err;
  return fr.status(pdeqx_conv_bwd_data(&d, fr.in(1), fr.in(2), fr.opt(3), nullptr, fr.out(0), fr.out(3), fr.out_bytes(3), stream),
                   "pdeqx_conv_bwd_data");
}

// ==========================================================================================
// PhysicsConvTranspose / MorePaddingConvTranspose (attrs as the conv + output_padding; no precision: fp32 kernels)
//   fwd operands: x, w, b or (0,)        results: y
//   bwd operands: x, gy, w               results: gx, gw, gb or (0,)
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_convT_fwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_convT_fwd", 3, 1);
  pdeqx_convT_desc d;
  if (!convT_desc(fr, fr.arg(0), fr.arg(1), &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: bad operand ranks or missing static attributes", "pdeqx_convT_fwd");
  if (!execute) return nullptr;
  return fr.status(pdeqx_convT_fwd(&d, fr.in(0), fr.in(1), fr.opt(2), fr.out(0), stream), "pdeqx_convT_fwd");
}

extern "C" XLA_FFI_Error* pdeqx_ffi_convT_bwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_convT_bwd", 3, 3);
  pdeqx_convT_desc d;
  if (!convT_desc(fr, fr.arg(0), fr.arg(2), &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: bad operand ranks or missing static attributes", "pdeqx_convT_bwd");
  if (!execute) return nullptr;
  if ((err = fr.status(pdeqx_convT_bwd_filter(&d, fr.in(0), fr.in(1), fr.out(1), fr.opt_out(2), stream), "pdeqx_convT_bwd_filter")))
    return err;
  return fr.status(pdeqx_convT_bwd_data(&d, fr.in(1), fr.in(2), fr.out(0), stream), "pdeqx_convT_bwd_data");
}

// ==========================================================================================
// PointwiseLinearConv
//   fwd operands: x [B,Ci,*S], w [Co,Ci,1..], b or (0,), add [B,Co,*S] or (0,)   attrs: activation   results: y
//   bwd operands: x, w, gy                                                       results: gx, gw, gb or (0,)
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_pointwise_fwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_pointwise_fwd", 4, 1);
  PDEQX_FFI_NUMBER(activation, "activation", "pdeqx_pointwise_fwd");
  const XLA_FFI_Buffer *x = fr.arg(0), *w = fr.arg(1);
  if (x->rank < 3 || w->rank < 2 || w->dims[1] != x->dims[1])
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: x must be [B, C_in, *S] and w [C_out, C_in, 1...]", "pdeqx_pointwise_fwd");
  if (!execute) return nullptr;
  return fr.status(pdeqx_pointwise_fwd((int32_t)x->dims[0], (int32_t)x->dims[1], (int32_t)w->dims[0], pixels(x), fr.in(0), fr.in(1),
                                       fr.opt(2), fr.opt(3), (int32_t)activation, fr.out(0), stream),
                   "pdeqx_pointwise_fwd");
}

extern "C" XLA_FFI_Error* pdeqx_ffi_pointwise_bwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_pointwise_bwd", 3, 3);
  const XLA_FFI_Buffer *x = fr.arg(0), *w = fr.arg(1);
  if (x->rank < 3 || w->rank < 2 || w->dims[1] != x->dims[1])
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: x must be [B, C_in, *S] and w [C_out, C_in, 1...]", "pdeqx_pointwise_bwd");
  if (!execute) return nullptr;
  return fr.status(pdeqx_pointwise_bwd((int32_t)x->dims[0], (int32_t)x->dims[1], (int32_t)w->dims[0], pixels(x), fr.in(0), fr.in(1),
                                       fr.in(2), fr.out(0), fr.out(1), fr.opt_out(2), stream),
                   "pdeqx_pointwise_bwd");
}

// ==========================================================================================
// GroupNorm
//   fwd operands: x [B,C,*S], weight [C] or (0,), bias [C] or (0,)   attrs: groups, eps
//       results : y, stats [B,groups,2] (saved), workspace (pdeqx_groupnorm_workspace_bytes)
//   bwd operands: x, weight or (0,), stats, gy                      attrs: groups
//       results : gx, gweight or (0,), gbias or (0,), scratch [2*B*C]
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_groupnorm_fwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_groupnorm_fwd", 3, 3);
  PDEQX_FFI_NUMBER(groups, "groups", "pdeqx_groupnorm_fwd");
  PDEQX_FFI_NUMBER(eps, "eps", "pdeqx_groupnorm_fwd");
  const XLA_FFI_Buffer* x = fr.arg(0);
  if (x->rank < 3) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: x must be [B, C, *S]", "pdeqx_groupnorm_fwd");
  if (fr.out_bytes(2) < pdeqx_groupnorm_workspace_bytes((int32_t)x->dims[0], (int32_t)x->dims[1]))
    return fr.fail(XLA_FFI_Error_Code_RESOURCE_EXHAUSTED, "%s: workspace result is smaller than pdeqx_groupnorm_workspace_bytes", "pdeqx_groupnorm_fwd");
  if (!execute) return nullptr;
  return fr.status(pdeqx_groupnorm_fwd((int32_t)x->dims[0], (int32_t)x->dims[1], (int32_t)groups, pixels(x), fr.in(0), fr.opt(1), fr.opt(2),
                                       (float)eps, fr.out(0), fr.out(1), fr.out(2), stream),
                   "pdeqx_groupnorm_fwd");
}

extern "C" XLA_FFI_Error* pdeqx_ffi_groupnorm_bwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_groupnorm_bwd", 4, 4);
  PDEQX_FFI_NUMBER(groups, "groups", "pdeqx_groupnorm_bwd");
  const XLA_FFI_Buffer* x = fr.arg(0);
  if (x->rank < 3) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: x must be [B, C, *S]", "pdeqx_groupnorm_bwd");
  if (Frame::count(fr.ret(3)) < 2 * x->dims[0] * x->dims[1])
    return fr.fail(XLA_FFI_Error_Code_RESOURCE_EXHAUSTED, "%s: scratch result must hold 2 * B * C floats", "pdeqx_groupnorm_bwd");
  if (!execute) return nullptr;
  return fr.status(pdeqx_groupnorm_bwd((int32_t)x->dims[0], (int32_t)x->dims[1], (int32_t)groups, pixels(x), fr.in(0), fr.opt(1), fr.in(2),
                                       fr.in(3), nullptr, fr.out(0), fr.opt_out(1), fr.opt_out(2), fr.out(3), stream),
                   "pdeqx_groupnorm_bwd");
}

// ==========================================================================================
// GroupNorm (groups = 1) -> PhysicsConv -> activation as ONE op, the norm folded into the convolution's passes
// (pdequinox/blocks/_dilated_res_block.py:141-151; accepted shapes: pdeqx_conv_gn_fusable)
//   fwd operands: x [B,C,*S], gn_weight [C] or (0,), gn_bias [C] or (0,), w [Co,C,*k], b [Co,1..] or (0,),
//                 in_parts [B,slots_in,2] or (0,): per-sample sums (x, x^2) its producer wrote (absent: a statistics pass)
//       attrs   : the conv's + activation, eps
//       results : y, stats [B,2] (saved), out_parts [B, pdeqx_conv_gn_slots, 2] (sums of y, y^2 for the next norm),
//                 workspace (max of pdeqx_conv_workspace_bytes and pdeqx_groupnorm_workspace_bytes)
//   bwd operands: x, stats, gn_weight or (0,), gn_bias or (0,), w, ga [B,Co,*S] (cotangent of the PRE-activation),
//                 add [B,C,*S] or (0,)                                   attrs: the conv's + relu_mask_x (0 / 1)
//       results : gx, g_gn_weight or (0,), g_gn_bias or (0,), gw, gb or (0,), v [B,C,*S] (scratch: cotangent of the
//                 normalised tensor), parts [B, pdeqx_conv_gn_slots, 2], scratch [2*B*C], workspace
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_norm_conv_fwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_norm_conv_fwd", 6, 4);
  PDEQX_FFI_NUMBER(activation, "activation", "pdeqx_norm_conv_fwd");
  PDEQX_FFI_NUMBER(eps, "eps", "pdeqx_norm_conv_fwd");
  pdeqx_conv_desc d;
  const XLA_FFI_Buffer* x = fr.arg(0);
  if (!conv_desc(fr, x, fr.arg(3), &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: bad operand ranks or missing static attributes", "pdeqx_norm_conv_fwd");
  const int slots = pdeqx_conv_gn_slots(&d);
  if (slots == 0)
    return fr.fail(XLA_FFI_Error_Code_UNIMPLEMENTED, "%s: this convolution does not take the folded GroupNorm passes (pdeqx_conv_gn_fusable)", "pdeqx_norm_conv_fwd");
  if (Frame::count(fr.ret(1)) < 2 * x->dims[0] || Frame::count(fr.ret(2)) < 2 * (int64_t)slots * x->dims[0])
    return fr.fail(XLA_FFI_Error_Code_RESOURCE_EXHAUSTED, "%s: stats / out_parts results are too small", "pdeqx_norm_conv_fwd");
  size_t need = pdeqx_conv_workspace_bytes(&d);
  const size_t gn_ws = pdeqx_groupnorm_workspace_bytes(d.batch, d.c_in);
  if (gn_ws > need) need = gn_ws;
  if (fr.out_bytes(3) < need)
    return fr.fail(XLA_FFI_Error_Code_RESOURCE_EXHAUSTED, "%s: workspace result is too small", "pdeqx_norm_conv_fwd");
  if (!execute) return nullptr;
  const XLA_FFI_Buffer* in_parts = fr.arg(5);
  if (Frame::count(in_parts) == 0) {
    if ((err = fr.status(pdeqx_groupnorm_stats(d.batch, d.c_in, 1, pixels(x), fr.in(0), (float)eps, fr.out(1), fr.out(3), stream),
                         "pdeqx_groupnorm_stats")))
      return err;
  } else {
    if (in_parts->rank != 3 || in_parts->dims[0] != x->dims[0] || in_parts->dims[2] != 2)
      return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: in_parts must be [B, slots, 2]", "pdeqx_norm_conv_fwd");
    if ((err = fr.status(pdeqx_groupnorm_stats_from_parts(d.batch, (int32_t)in_parts->dims[1], (int64_t)d.c_in * pixels(x), fr.in(5),
                                                          (float)eps, fr.out(1), stream),
                         "pdeqx_groupnorm_stats_from_parts")))
      return err;
  }
  return fr.status(pdeqx_conv_gn_fwd(&d, fr.in(0), fr.in(3), fr.opt(4), fr.out(1), fr.opt(1), fr.opt(2), (int32_t)activation, fr.out(0),
                                     fr.out(2), fr.out(3), fr.out_bytes(3), stream),
                   "pdeqx_conv_gn_fwd");
}

extern "C" XLA_FFI_Error* pdeqx_ffi_norm_conv_bwd(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_norm_conv_bwd", 7, 9);
  PDEQX_FFI_NUMBER(relu_mask_x, "relu_mask_x", "pdeqx_norm_conv_bwd");
  pdeqx_conv_desc d;
  const XLA_FFI_Buffer* x = fr.arg(0);
  if (!conv_desc(fr, x, fr.arg(4), &d))
    return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: bad operand ranks or missing static attributes", "pdeqx_norm_conv_bwd");
  const int slots = pdeqx_conv_gn_slots(&d);
  if (slots == 0)
    return fr.fail(XLA_FFI_Error_Code_UNIMPLEMENTED, "%s: this convolution does not take the folded GroupNorm passes (pdeqx_conv_gn_fusable)", "pdeqx_norm_conv_bwd");
  if (Frame::count(fr.ret(5)) < Frame::count(x) || Frame::count(fr.ret(6)) < 2 * (int64_t)slots * x->dims[0] ||
      Frame::count(fr.ret(7)) < 2 * x->dims[0] * x->dims[1] || fr.out_bytes(8) < pdeqx_conv_workspace_bytes(&d))
    return fr.fail(XLA_FFI_Error_Code_RESOURCE_EXHAUSTED, "%s: a scratch result (v, parts, scratch, workspace) is too small", "pdeqx_norm_conv_bwd");
  if (!execute) return nullptr;
  if ((err = fr.status(pdeqx_conv_gn_bwd_filter(&d, fr.in(0), fr.in(5), fr.in(1), fr.opt(2), fr.opt(3), fr.out(3), fr.opt_out(4), fr.out(8),
                                                fr.out_bytes(8), stream),
                       "pdeqx_conv_gn_bwd_filter")))
    return err;
  if ((err = fr.status(pdeqx_conv_gn_bwd_data(&d, fr.in(5), fr.in(4), fr.opt(2), fr.in(0), fr.out(5), fr.out(6), fr.out(8), fr.out_bytes(8), stream),
                       "pdeqx_conv_gn_bwd_data")))
    return err;
  return fr.status(pdeqx_groupnorm_bwd_fused(d.batch, d.c_in, pixels(x), fr.in(0), fr.out(5), fr.opt(2), fr.in(1), fr.out(6), slots,
                                             relu_mask_x != 0 ? 1 : 0, fr.opt(6), fr.out(0), fr.opt_out(1), fr.opt_out(2), fr.out(7), stream),
                   "pdeqx_groupnorm_bwd_fused");
}

// ==========================================================================================
// Adam (optax.adam + eqx.apply_updates on one flat leaf), IN PLACE:
//   operands: param, grad, mu, nu, state [4] (state[0] = update count, int32 bit pattern, starts at 0)
//   attrs   : lr, b1, b2, eps, grad_scale (1 / number of data-parallel ranks)
//   results : param, mu, nu, state -- each must ALIAS its operand (jax.ffi.ffi_call(..., input_output_aliases={0:0, 2:1, 3:2, 4:3}))
// ==========================================================================================
extern "C" XLA_FFI_Error* pdeqx_ffi_adam(XLA_FFI_CallFrame* call_frame) {
  PDEQX_FFI_PROLOGUE("pdeqx_adam", 5, 4);
  PDEQX_FFI_NUMBER(lr, "lr", "pdeqx_adam");
  PDEQX_FFI_NUMBER(b1, "b1", "pdeqx_adam");
  PDEQX_FFI_NUMBER(b2, "b2", "pdeqx_adam");
  PDEQX_FFI_NUMBER(eps, "eps", "pdeqx_adam");
  PDEQX_FFI_NUMBER(grad_scale, "grad_scale", "pdeqx_adam");
  const int alias[4] = {0, 2, 3, 4};
  for (int r = 0; r < 4; ++r)
    if (fr.ret(r)->data != fr.arg(alias[r])->data)
      return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: result %lld must alias its operand (input_output_aliases)", "pdeqx_adam", r);
  if (Frame::count(fr.arg(4)) < 4) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "%s: state must hold 4 floats", "pdeqx_adam");
  if (!execute) return nullptr;
  return fr.status(pdeqx_adam_step_dev(Frame::count(fr.arg(0)), fr.out(0), fr.in(1), fr.out(1), fr.out(2), (float)lr, (float)b1,
                                       (float)b2, (float)eps, fr.out(3), (float)grad_scale, stream),
                   "pdeqx_adam_step_dev");
}

// names of every handler symbol, for registration loops and the export test
extern "C" const char* const* pdeqx_ffi_handler_names(int* count) {
  static const char* const names[] = {"pdeqx_ffi_spectral_block_fwd", "pdeqx_ffi_spectral_block_bwd", "pdeqx_ffi_conv_fwd",
                                      "pdeqx_ffi_conv_bwd",           "pdeqx_ffi_convT_fwd",          "pdeqx_ffi_convT_bwd",
                                      "pdeqx_ffi_pointwise_fwd",      "pdeqx_ffi_pointwise_bwd",      "pdeqx_ffi_groupnorm_fwd",
                                      "pdeqx_ffi_groupnorm_bwd",      "pdeqx_ffi_adam",               "pdeqx_ffi_norm_conv_fwd",
                                      "pdeqx_ffi_norm_conv_bwd"};
  if (count) *count = (int)(sizeof(names) / sizeof(names[0]));
  return names;
}
