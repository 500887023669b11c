# Builds the C-ABI CUDA library in-tree (sm_100a only) and the oracle's C restatement.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude --fmad=false
LIB := xenoverse_b200/libamdp_b200.so

all: $(LIB)

$(LIB): xenoverse_b200/csrc/amdp_b200.cu include/amdp_b200.h
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o $@ $< 2> build_ptxas.log || (cat build_ptxas.log; false)

clean:
	rm -f $(LIB) build_ptxas.log
