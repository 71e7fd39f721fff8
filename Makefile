# Builds libpdeqx_b200.so (sm_100a only) in-tree so that it travels with the gpurun snapshot.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
SRC := $(wildcard pdequinox_b200/csrc/*.cu)
OBJ := $(patsubst pdequinox_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := pdequinox_b200/libpdeqx_b200.so

all: $(LIB)

build/%.o: pdequinox_b200/csrc/%.cu $(wildcard pdequinox_b200/csrc/*.h pdequinox_b200/csrc/*.cuh) include/pdeqx_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -o $@ $^

clean:
	rm -rf build $(LIB)
.PHONY: all clean
