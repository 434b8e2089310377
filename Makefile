# Builds the CUDA library (sm_100a only) and the CPU oracle.
PKG := plutus-halo2-verifier-gen_b200
NVCC ?= /usr/local/cuda/bin/nvcc
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper); name the system compiler for nvcc explicitly
HOSTCXX ?= /usr/bin/g++
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin $(HOSTCXX) \
           -Iinclude -I$(PKG)/csrc
SRCS := $(PKG)/csrc/b200zk.cu $(PKG)/csrc/b200zk_ext.cu $(PKG)/csrc/b200zk_diag.cu $(PKG)/csrc/h2mo.cu
OBJS := $(SRCS:$(PKG)/csrc/%.cu=build/%.o)
HDRS := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.hpp) include/b200zk.h

all:
	$(MAKE) -j4 lib
	$(MAKE) oracle

lib: $(PKG)/libb200zk.so

# four translation units (MSM + NTT + contexts; polynomial side + SRS + decompression; self-test + micro-benchmarks;
# transcript + multi-open + guards) compiled side by side
build/%.o: $(PKG)/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(PKG)/libb200zk.so: $(OBJS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(OBJS) -lcudart

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(PKG)/libb200zk.so build
	$(MAKE) -C oracle clean
.PHONY: all lib oracle clean
