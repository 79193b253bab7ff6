# Builds libb2f.so (hand-written sm_100a kernels + C ABI) in-tree.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -cudart shared \
             --expt-relaxed-constexpr -Xptxas -v
CSRC      := gpt_image_edit_b200/csrc
LIBDIR    := gpt_image_edit_b200/lib
# attention_experiments.cu (alternative kernel structures kept for the record) is not part of
# the product library: `make EXPERIMENTS=1` links it and enables B2F_ATTN_VARIANT 10-12 / 30-32 / 40-42 / 60-62.
SRCS      := $(filter-out $(CSRC)/attention_experiments.cu,$(wildcard $(CSRC)/*.cu))
ifeq ($(EXPERIMENTS),1)
SRCS      += $(CSRC)/attention_experiments.cu
NVFLAGS   += -DB2F_WITH_EXPERIMENTS
endif
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/b2f.h

all: $(LIBDIR)/libb2f.so

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)
	@grep -E "error|warning|spill|registers" build/$*.ptxas.log | grep -v "0 bytes spill" | head -40 || true

$(LIBDIR)/libb2f.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -cudart shared -o $@ $(OBJS) -Xlinker -rpath -Xlinker /usr/local/cuda/lib64

clean:
	rm -rf build $(LIBDIR)/libb2f.so

.PHONY: all clean
