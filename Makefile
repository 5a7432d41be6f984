# Builds libsldm_sage.so (the C-ABI CUDA library, sm_100a only) and the CPU oracle.
# `python __graft_entry__.py` (build()) drives this; `make` works on its own too.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function \
             -Xptxas -warn-spills --expt-relaxed-constexpr
CSRC      := sldm_gnn_b200/csrc
OBJDIR    := build/obj
LIB       := sldm_gnn_b200/lib/libsldm_sage.so
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) include/sldm_sage.h

all: $(LIB) oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -cudart static

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
