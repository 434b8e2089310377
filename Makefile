# Builds the CUDA library (sm_100a only) and the CPU oracle.
PKG := plutus-halo2-verifier-gen_b200
NVCC ?= /usr/local/cuda/bin/nvcc
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper); name the system compiler for nvcc explicitly
HOSTCXX ?= /usr/bin/g++
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin $(HOSTCXX) \
           -Iinclude -I$(PKG)/csrc
SRCS := $(PKG)/csrc/b200zk.cu
HDRS := $(wildcard $(PKG)/csrc/*.cuh) include/b200zk.h

all: $(PKG)/libb200zk.so oracle

$(PKG)/libb200zk.so: $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS) -lcudart

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(PKG)/libb200zk.so
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
