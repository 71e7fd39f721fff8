/* MINIMAL RESTATEMENT of the XLA FFI C API (xla/ffi/api/c_api.h, API version 0.1) -- COMPILE-CHECK ONLY.
 *
 * jax / jaxlib are not installed in the build image of this repository, so the real header is not available here.
 * pdeqx_xla_ffi.cc uses nothing but the plain-C, struct_size-versioned call-frame ABI; this file restates the few
 * structs it touches so that the shim can be compiled, loaded and driven by tests (tests/test_xla_ffi_shim*.py build
 * call frames by hand). On a machine with JAX the Makefile compiles the shim against the REAL header
 * (-I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")) and this directory is not on the include path.
 * Field order follows the upstream header; anything the shim does not read is left out of the trailing part of
 * XLA_FFI_Api only (a struct it never allocates, only reads through a pointer, leading fields first).
 */
#ifndef PDEQX_XLA_FFI_C_API_STUB_H_
#define PDEQX_XLA_FFI_C_API_STUB_H_

#include <stddef.h>
#include <stdint.h>

#define PDEQX_XLA_FFI_STUB_HEADER 1
#define XLA_FFI_API_MAJOR 0
#define XLA_FFI_API_MINOR 1

#ifdef __cplusplus
extern "C" {
#endif

typedef struct XLA_FFI_Api XLA_FFI_Api;
typedef struct XLA_FFI_InternalApi XLA_FFI_InternalApi;
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_ExecutionContext XLA_FFI_ExecutionContext;
typedef struct XLA_FFI_Future XLA_FFI_Future;

typedef enum { XLA_FFI_Extension_Metadata = 1 } XLA_FFI_Extension_Type;

typedef struct XLA_FFI_Extension_Base {
  size_t struct_size;
  XLA_FFI_Extension_Type type;
  struct XLA_FFI_Extension_Base* next;
} XLA_FFI_Extension_Base;

typedef struct XLA_FFI_Api_Version {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int major_version; /* out */
  int minor_version; /* out */
} XLA_FFI_Api_Version;

typedef enum {
  XLA_FFI_Error_Code_OK = 0,
  XLA_FFI_Error_Code_CANCELLED = 1,
  XLA_FFI_Error_Code_UNKNOWN = 2,
  XLA_FFI_Error_Code_INVALID_ARGUMENT = 3,
  XLA_FFI_Error_Code_DEADLINE_EXCEEDED = 4,
  XLA_FFI_Error_Code_NOT_FOUND = 5,
  XLA_FFI_Error_Code_ALREADY_EXISTS = 6,
  XLA_FFI_Error_Code_PERMISSION_DENIED = 7,
  XLA_FFI_Error_Code_RESOURCE_EXHAUSTED = 8,
  XLA_FFI_Error_Code_FAILED_PRECONDITION = 9,
  XLA_FFI_Error_Code_ABORTED = 10,
  XLA_FFI_Error_Code_OUT_OF_RANGE = 11,
  XLA_FFI_Error_Code_UNIMPLEMENTED = 12,
  XLA_FFI_Error_Code_INTERNAL = 13,
  XLA_FFI_Error_Code_UNAVAILABLE = 14,
  XLA_FFI_Error_Code_DATA_LOSS = 15,
  XLA_FFI_Error_Code_UNAUTHENTICATED = 16
} XLA_FFI_Error_Code;

typedef struct XLA_FFI_Error_Create_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const char* message;
  XLA_FFI_Error_Code errc;
} XLA_FFI_Error_Create_Args;
typedef XLA_FFI_Error* XLA_FFI_Error_Create(XLA_FFI_Error_Create_Args* args);

typedef enum {
  XLA_FFI_DataType_INVALID = 0,
  XLA_FFI_DataType_PRED = 1,
  XLA_FFI_DataType_S8 = 2,
  XLA_FFI_DataType_S16 = 3,
  XLA_FFI_DataType_S32 = 4,
  XLA_FFI_DataType_S64 = 5,
  XLA_FFI_DataType_U8 = 6,
  XLA_FFI_DataType_U16 = 7,
  XLA_FFI_DataType_U32 = 8,
  XLA_FFI_DataType_U64 = 9,
  XLA_FFI_DataType_F16 = 10,
  XLA_FFI_DataType_F32 = 11,
  XLA_FFI_DataType_F64 = 12,
  XLA_FFI_DataType_C64 = 15,
  XLA_FFI_DataType_BF16 = 16,
  XLA_FFI_DataType_TOKEN = 17,
  XLA_FFI_DataType_C128 = 18
} XLA_FFI_DataType;

typedef struct XLA_FFI_Buffer {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_DataType dtype;
  void* data;
  int64_t* dims;
  int64_t rank;
} XLA_FFI_Buffer;

typedef enum { XLA_FFI_ArgType_BUFFER = 1 } XLA_FFI_ArgType;
typedef enum { XLA_FFI_RetType_BUFFER = 1 } XLA_FFI_RetType;
typedef enum {
  XLA_FFI_AttrType_ARRAY = 1,
  XLA_FFI_AttrType_DICTIONARY = 2,
  XLA_FFI_AttrType_SCALAR = 3,
  XLA_FFI_AttrType_STRING = 4
} XLA_FFI_AttrType;

typedef enum {
  XLA_FFI_ExecutionStage_INSTANTIATE = 0,
  XLA_FFI_ExecutionStage_PREPARE = 1,
  XLA_FFI_ExecutionStage_INITIALIZE = 2,
  XLA_FFI_ExecutionStage_EXECUTE = 3
} XLA_FFI_ExecutionStage;

typedef struct XLA_FFI_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_ArgType* types; /* length == size */
  void** args;            /* length == size */
} XLA_FFI_Args;

typedef struct XLA_FFI_Rets {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_RetType* types; /* length == size */
  void** rets;            /* length == size */
} XLA_FFI_Rets;

typedef struct XLA_FFI_ByteSpan {
  const char* ptr;
  size_t len;
} XLA_FFI_ByteSpan;

typedef struct XLA_FFI_Scalar {
  XLA_FFI_DataType dtype;
  void* value;
} XLA_FFI_Scalar;

typedef struct XLA_FFI_Array {
  XLA_FFI_DataType dtype;
  size_t size;
  void* data;
} XLA_FFI_Array;

typedef struct XLA_FFI_Attrs {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_AttrType* types;  /* length == size */
  XLA_FFI_ByteSpan** names; /* length == size */
  void** attributes;        /* length == size */
} XLA_FFI_Attrs;

typedef struct XLA_FFI_CallFrame {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const XLA_FFI_Api* api;
  XLA_FFI_ExecutionContext* ctx;
  XLA_FFI_ExecutionStage stage;
  XLA_FFI_Args args;
  XLA_FFI_Rets rets;
  XLA_FFI_Attrs attrs;
  XLA_FFI_Future* future; /* out */
} XLA_FFI_CallFrame;

typedef XLA_FFI_Error* XLA_FFI_Handler(XLA_FFI_CallFrame* call_frame);

typedef uint32_t XLA_FFI_Handler_Traits;
enum XLA_FFI_Handler_TraitsBits { XLA_FFI_HANDLER_TRAITS_COMMAND_BUFFER_COMPATIBLE = 1u << 0 };

typedef struct XLA_FFI_Metadata {
  size_t struct_size;
  XLA_FFI_Api_Version api_version;
  XLA_FFI_Handler_Traits traits;
} XLA_FFI_Metadata;

typedef struct XLA_FFI_Metadata_Extension {
  XLA_FFI_Extension_Base extension_base;
  XLA_FFI_Metadata* metadata;
} XLA_FFI_Metadata_Extension;

typedef struct XLA_FFI_Error_GetMessage_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_Error* error;
  const char* message; /* out */
} XLA_FFI_Error_GetMessage_Args;
typedef void XLA_FFI_Error_GetMessage(XLA_FFI_Error_GetMessage_Args* args);

typedef struct XLA_FFI_Error_Destroy_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_Error* error;
} XLA_FFI_Error_Destroy_Args;
typedef void XLA_FFI_Error_Destroy(XLA_FFI_Error_Destroy_Args* args);

typedef struct XLA_FFI_Handler_Register_Args XLA_FFI_Handler_Register_Args;
typedef XLA_FFI_Error* XLA_FFI_Handler_Register(XLA_FFI_Handler_Register_Args* args);

typedef struct XLA_FFI_Stream_Get_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_ExecutionContext* ctx;
  void* stream; /* out */
} XLA_FFI_Stream_Get_Args;
typedef XLA_FFI_Error* XLA_FFI_Stream_Get(XLA_FFI_Stream_Get_Args* args);

/* leading part of the function table (the shim reads XLA_FFI_Error_Create and XLA_FFI_Stream_Get only) */
struct XLA_FFI_Api {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_Api_Version api_version;
  XLA_FFI_InternalApi* internal_api;
  XLA_FFI_Error_Create* XLA_FFI_Error_Create;
  XLA_FFI_Error_GetMessage* XLA_FFI_Error_GetMessage;
  XLA_FFI_Error_Destroy* XLA_FFI_Error_Destroy;
  XLA_FFI_Handler_Register* XLA_FFI_Handler_Register;
  XLA_FFI_Stream_Get* XLA_FFI_Stream_Get;
  /* ... the upstream table continues (type ids, execution context, state, device memory, thread pool, futures) */
};

#ifdef __cplusplus
}
#endif
#endif /* PDEQX_XLA_FFI_C_API_STUB_H_ */
