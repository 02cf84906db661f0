# Builds libmtlora_b200.so (hand-written sm_100a kernels behind the C ABI of include/mtlora_b200.h) in-tree,
# and the CPU oracle helpers. nvcc cross-compiles without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
CSRC      := mtlora_b200/csrc
OBJS      := $(CSRC)/api.o $(CSRC)/linear_sm100.o $(CSRC)/attention.o $(CSRC)/attention_sm100.o $(CSRC)/rowwise.o $(CSRC)/xty.o $(CSRC)/xty_sm100.o $(CSRC)/patch_embed.o $(CSRC)/optim.o $(CSRC)/rankproj.o
LIB       := mtlora_b200/libmtlora_b200.so

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/kernels.cuh $(CSRC)/linear_sm100.cuh include/mtlora_b200.h
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB)

.PHONY: all clean
