# Build libscirs2_fft_cuda.so (sm_100a only) and the CPU oracle.
NVCC      ?= nvcc
EXTRA     ?=
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
             --expt-relaxed-constexpr -Xptxas -v $(EXTRA)
CSRC      := scirs_b200/csrc
BUILD     := build
LIBDIR    := scirs_b200/lib
LIB       := $(LIBDIR)/libscirs2_fft_cuda.so

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS))
HDRS      := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/scirs2_fft_cuda.h

all: $(LIB) oracle

$(BUILD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -Xlinker --version-script=$(CSRC)/exports.map

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) $(LIB)

.PHONY: all oracle clean
