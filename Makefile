# Builds the C-ABI shared library (CUDA, sm_100a).  `python -c "import __graft_entry__ as g; g.build()"` calls this.
# (The oracle is numpy / scipy: nothing to compile for it; the reference is pure Python: there is no oracle/_ref.)
# -fmad=false: no FMA contraction, so every kernel family (interpreter, term list, Laplacian kernels) rounds like the
# numpy reference and like each other -- partitioned and one-GPU runs stay bit-identical whatever kernel a row lands in.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := gds_b200/csrc
LIB  := gds_b200/lib/libgds_b200.so
SRCS := $(CSRC)/kernels.cu $(CSRC)/api.cu $(CSRC)/lower.cpp $(CSRC)/lattice.cpp
HDRS := $(CSRC)/internal.h include/gds_b200.h
NVFLAGS := -O3 -std=c++17 -lineinfo -fmad=false $(EXTRA) $(ARCH) -Iinclude -I$(CSRC) -Xcompiler -fPIC,-O3,-Wall -shared -cudart shared

all: $(LIB)

$(LIB): $(SRCS) $(HDRS) Makefile
	@mkdir -p gds_b200/lib
	$(NVCC) $(NVFLAGS) -Xptxas -v $(SRCS) -o $(LIB) 2> gds_b200/lib/ptxas.log || (cat gds_b200/lib/ptxas.log; exit 1)
	@grep -E "registers|spill" gds_b200/lib/ptxas.log | sort | uniq -c | head -20

clean:
	rm -rf gds_b200/lib
.PHONY: all clean
