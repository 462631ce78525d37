# Builds libdgvit.so (sm_100a only) and nothing else.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
PKG   := dgvit-depth-goal-guided-vision-transformer-_b200
CSRC  := $(PKG)/csrc
NVCC  ?= /usr/local/cuda/bin/nvcc
FLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall \
         -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
LIB   := $(PKG)/libdgvit.so
SRCS  := $(CSRC)/dgvit.cu $(CSRC)/depth.cu
HDRS  := $(wildcard $(CSRC)/*.cuh) include/dgvit.h

all: $(LIB)

OBJS  := $(PKG)/csrc/.obj/dgvit.o $(PKG)/csrc/.obj/depth.o

$(PKG)/csrc/.obj/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(PKG)/csrc/.obj
	$(NVCC) $(FLAGS) $(EXTRA) -DDGVIT_WITH_TC -c -o $@ $<

$(LIB): $(OBJS)
	$(NVCC) $(FLAGS) -shared -o $@ $(OBJS)

# instrumented build for profiles/mlp_trace.py (never loaded by the package)
trace: $(SRCS) $(HDRS)
	$(NVCC) $(FLAGS) -DDGVIT_WITH_TC -DDGVIT_MLP_TRACE -shared -o $(PKG)/libdgvit_trace.so $(SRCS)

clean:
	rm -rf $(LIB) $(PKG)/csrc/.obj
.PHONY: all clean trace
