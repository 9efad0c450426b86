# Builds libmmif_b200.so (sm_100a only) in-tree, plus nothing else: the oracle is Python.
NVCC ?= nvcc
PKG := multi-modal-image-fusion_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/mmif_b200.h
OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
LIB := $(PKG)/libmmif_b200.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 \
           -Xptxas -v --expt-relaxed-constexpr

all: $(LIB)

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
	@grep -E "error|warning" build/$*.ptxas.log | grep -v "Compiling entry" || true

$(LIB): $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ) -cudart static

clean:
	rm -rf build $(LIB)
.PHONY: all clean
