# Builds libpdeqx_b200.so (sm_100a only) in-tree so that it travels with the gpurun snapshot.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
SRC := $(wildcard pdequinox_b200/csrc/*.cu)
OBJ := $(patsubst pdequinox_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := pdequinox_b200/libpdeqx_b200.so

FFI_LIB := pdequinox_b200/libpdeqx_xla_ffi.so
# XLA FFI shim: against jaxlib's own xla/ffi/api/c_api.h when JAX is importable, else against the minimal restatement of
# that header under csrc/xla_ffi/stub (compile check + the call-frame tests; JAX is not installable in this image)
FFI_INC := $(shell python -c "import jax.ffi; print(jax.ffi.include_dir())" 2>/dev/null || echo pdequinox_b200/csrc/xla_ffi/stub)

all: $(LIB) $(FFI_LIB)

xla_ffi: $(FFI_LIB)

$(FFI_LIB): pdequinox_b200/csrc/xla_ffi/pdeqx_xla_ffi.cc include/pdeqx_b200.h $(LIB)
	g++ -O2 -std=c++17 -fPIC -shared -Wall -Wno-format-security -I$(FFI_INC) -Iinclude -o $@ $< -Lpdequinox_b200 -lpdeqx_b200 -Wl,-rpath,'$$ORIGIN'

build/%.o: pdequinox_b200/csrc/%.cu $(wildcard pdequinox_b200/csrc/*.h pdequinox_b200/csrc/*.cuh) include/pdeqx_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -o $@ $^

clean:
	rm -rf build $(LIB) $(FFI_LIB)
.PHONY: all clean xla_ffi
