# Top-level build: libsalmon_b200.so (CUDA, sm_100a only) + the oracle (test infra).
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -fopenmp $(PTXAS_V)
CSRC      := salmon_b200/csrc
LIB       := salmon_b200/libsalmon_b200.so
SRCS      := $(wildcard $(CSRC)/*.cu)
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: $(LIB) oracle

$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRCS) -ldl -lgomp -lz

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB); $(MAKE) -C oracle clean

.PHONY: all oracle clean
