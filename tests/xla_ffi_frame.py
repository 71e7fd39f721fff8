"""Hand-built XLA FFI call frames (ctypes mirror of xla/ffi/api/c_api.h as restated in
pdequinox_b200/csrc/xla_ffi/stub) to drive libpdeqx_xla_ffi.so without JAX: what XLA's runtime does when it executes a
`jax.ffi.ffi_call` -- operands / results as XLA_FFI_Buffer, static attributes as scalars / arrays by name, the stream
through XLA_FFI_Stream_Get, errors through XLA_FFI_Error_Create."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "pdequinox_b200", "libpdeqx_xla_ffi.so")

F32, S32, S64, F64 = 11, 4, 5, 12
ARG_BUFFER = RET_BUFFER = 1
ATTR_ARRAY, ATTR_SCALAR = 1, 3
STAGE_INITIALIZE, STAGE_EXECUTE = 2, 3
EXT_METADATA = 1


class ExtBase(C.Structure):
    pass


ExtBase._fields_ = [("struct_size", C.c_size_t), ("type", C.c_int), ("next", C.POINTER(ExtBase))]


class ApiVersion(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("major_version", C.c_int), ("minor_version", C.c_int)]


class Buffer(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("dtype", C.c_int), ("data", C.c_void_p),
                ("dims", C.POINTER(C.c_int64)), ("rank", C.c_int64)]


class Args(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64), ("types", C.POINTER(C.c_int)),
                ("args", C.POINTER(C.c_void_p))]


class ByteSpan(C.Structure):
    _fields_ = [("ptr", C.c_char_p), ("len", C.c_size_t)]


class Scalar(C.Structure):
    _fields_ = [("dtype", C.c_int), ("value", C.c_void_p)]


class Array(C.Structure):
    _fields_ = [("dtype", C.c_int), ("size", C.c_size_t), ("data", C.c_void_p)]


class Attrs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64), ("types", C.POINTER(C.c_int)),
                ("names", C.POINTER(C.POINTER(ByteSpan))), ("attributes", C.POINTER(C.c_void_p))]


class ErrorCreateArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("message", C.c_char_p), ("errc", C.c_int)]


class StreamGetArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("ctx", C.c_void_p), ("stream", C.c_void_p)]


ERROR_CREATE = C.CFUNCTYPE(C.c_void_p, C.POINTER(ErrorCreateArgs))
STREAM_GET = C.CFUNCTYPE(C.c_void_p, C.POINTER(StreamGetArgs))


class Api(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("api_version", ApiVersion), ("internal_api", C.c_void_p),
                ("XLA_FFI_Error_Create", ERROR_CREATE), ("XLA_FFI_Error_GetMessage", C.c_void_p), ("XLA_FFI_Error_Destroy", C.c_void_p),
                ("XLA_FFI_Handler_Register", C.c_void_p), ("XLA_FFI_Stream_Get", STREAM_GET)]


class CallFrame(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.POINTER(ExtBase)), ("api", C.POINTER(Api)), ("ctx", C.c_void_p),
                ("stage", C.c_int), ("args", Args), ("rets", Args), ("attrs", Attrs), ("future", C.c_void_p)]


class Metadata(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("api_version", ApiVersion), ("traits", C.c_uint32)]


class MetadataExt(C.Structure):
    _fields_ = [("extension_base", ExtBase), ("metadata", C.POINTER(Metadata))]


class FfiError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"XLA_FFI_Error code {code}: {message}")
        self.code, self.message = code, message


def load_shim():
    lib = C.CDLL(SHIM)
    n = C.c_int()
    lib.pdeqx_ffi_handler_names.restype = C.POINTER(C.c_char_p)
    names = lib.pdeqx_ffi_handler_names(C.byref(n))
    handlers = {}
    for i in range(n.value):
        fn = getattr(lib, names[i].decode())
        fn.restype = C.c_void_p
        fn.argtypes = [C.POINTER(CallFrame)]
        handlers[names[i].decode()] = fn
    return lib, handlers


def query_metadata(handler):
    md = Metadata()
    md.struct_size = C.sizeof(Metadata)
    ext = MetadataExt()
    ext.extension_base.struct_size = C.sizeof(MetadataExt)
    ext.extension_base.type = EXT_METADATA
    ext.metadata = C.pointer(md)
    frame = CallFrame()
    frame.struct_size = C.sizeof(CallFrame)
    frame.extension_start = C.cast(C.pointer(ext), C.POINTER(ExtBase))
    rc = handler(C.byref(frame))
    return rc, md.api_version.major_version, md.api_version.minor_version, md.traits


class Caller:
    """Builds a call frame from (pointer, shape) operands / results and Python attribute values, and runs a handler."""

    def __init__(self, stream=0):
        self.errors = []
        self._keep = []

        def error_create(p):
            self.errors.append((p.contents.errc, p.contents.message.decode()))
            return 0xDEAD0  # any non-null XLA_FFI_Error*

        def stream_get(p):
            p.contents.stream = stream
            return None

        self._cbs = (ERROR_CREATE(error_create), STREAM_GET(stream_get))
        self.api = Api()
        self.api.struct_size = C.sizeof(Api)
        self.api.api_version.major_version, self.api.api_version.minor_version = 0, 1
        self.api.XLA_FFI_Error_Create, self.api.XLA_FFI_Stream_Get = self._cbs

    def _buffers(self, items):
        n = len(items)
        bufs = (Buffer * n)()
        ptrs = (C.c_void_p * n)()
        types = (C.c_int * n)()
        for i, (ptr, shape) in enumerate(items):
            dims = (C.c_int64 * max(len(shape), 1))(*shape)
            self._keep.append(dims)
            bufs[i].struct_size = C.sizeof(Buffer)
            bufs[i].dtype, bufs[i].data, bufs[i].dims, bufs[i].rank = F32, ptr, dims, len(shape)
            ptrs[i] = C.addressof(bufs[i])
            types[i] = ARG_BUFFER
        self._keep += [bufs, ptrs, types]
        a = Args()
        a.struct_size, a.size, a.types, a.args = C.sizeof(Args), n, types, ptrs
        return a

    def _attrs(self, attrs):
        n = len(attrs)
        types = (C.c_int * n)()
        names = (C.POINTER(ByteSpan) * n)()
        vals = (C.c_void_p * n)()
        for i, (k, v) in enumerate(attrs.items()):
            span = ByteSpan(k.encode(), len(k))
            names[i] = C.pointer(span)
            if isinstance(v, (list, tuple, np.ndarray)):
                arr = np.ascontiguousarray(v, dtype=np.int64)
                holder = Array(S64, arr.size, arr.ctypes.data)
                types[i] = ATTR_ARRAY
                self._keep.append(arr)
            else:
                val = np.array(v, dtype=np.float32 if isinstance(v, float) else np.int32)  # what jax.ffi sends for np scalars
                holder = Scalar(F32 if isinstance(v, float) else S32, val.ctypes.data)
                types[i] = ATTR_SCALAR
                self._keep.append(val)
            vals[i] = C.addressof(holder)
            self._keep += [span, holder]
        self._keep += [types, names, vals]
        a = Attrs()
        a.struct_size, a.size, a.types, a.names, a.attributes = C.sizeof(Attrs), n, types, names, vals
        return a

    def __call__(self, handler, args, rets, attrs, stage=STAGE_EXECUTE):
        """args / rets: lists of (device pointer or None, shape); attrs: dict. Raises FfiError if the handler reports one."""
        frame = CallFrame()
        frame.struct_size = C.sizeof(CallFrame)
        frame.api = C.pointer(self.api)
        frame.stage = stage
        frame.args = self._buffers(args)
        frame.rets = self._buffers(rets)
        frame.attrs = self._attrs(attrs)
        self.errors.clear()
        rc = handler(C.byref(frame))
        if rc:
            raise FfiError(*self.errors[-1])
