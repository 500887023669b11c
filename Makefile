# Builds the C-ABI CUDA library in-tree (sm_100a only).  One object per translation unit so that `make -j` compiles the
# rollout kernels in parallel; ptxas -v output is collected in build_ptxas.log.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude --fmad=false $(EXTRA)
SRC := xenoverse_b200/csrc
OBJ := build/amdp_b200.o build/amdp_mdp.o build/amdp_w.o build/amdp_tp.o build/amdp_vi.o build/amdp_sampler.o build/amdp_mt.o
LIB := xenoverse_b200/libamdp_b200.so

all: $(LIB)

build/%.o: $(SRC)/%.cu $(SRC)/amdp_dev.cuh include/amdp_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o $@ $< 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)
	@cat build/*.ptxas.log > build_ptxas.log

clean:
	rm -rf build $(LIB) build_ptxas.log
