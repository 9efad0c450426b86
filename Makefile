# Builds libmmif_b200.so (sm_100a only) in-tree, plus nothing else: the oracle is Python.
NVCC ?= nvcc
PKG := multi-modal-image-fusion_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/mmif_b200.h
OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
LIB := $(PKG)/libmmif_b200.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 \
           -Xptxas -v --expt-relaxed-constexpr $(EXTRA)

PYINC := $(shell python3 -c "import sysconfig; print(sysconfig.get_paths()['include'])")
FAST := $(PKG)/_fastcall.so

all: $(LIB) $(FAST)

# CPython fast-call binding of the per-step entries (optional: _lib.py falls back to ctypes without it)
$(FAST): $(PKG)/csrc/fastcall.c include/mmif_b200.h $(LIB)
	gcc -O2 -fPIC -shared -I$(PYINC) $< -o $@ -L$(PKG) -lmmif_b200 -Wl,-rpath,'$$ORIGIN'

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
	@grep -E "error|warning" build/$*.ptxas.log | grep -v "Compiling entry" || true

$(LIB): $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ) -cudart static

clean:
	rm -rf build $(LIB) $(FAST)
.PHONY: all clean
