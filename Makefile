# Builds the product library (libvrt.so: sm_100a kernels + C ABI), the host-side mirror (libvrt_host.so)
# and the CPU oracle (test infrastructure).  Everything is built in-tree so it travels with the repo snapshot.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
PKG := zig_vulkan_b200
CSRC := $(PKG)/csrc

# --fmad=false is part of the numerical contract (DESIGN.md "FP discipline"), not a tuning knob.
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -lineinfo \
             -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
HOSTFLAGS := -O2 -std=c++17 -fPIC -Wall -Wextra -pthread

KERNEL_HDRS := $(wildcard $(CSRC)/*.cuh) include/vrt.h

all: $(PKG)/libvrt.so $(PKG)/libvrt_host.so oracle

$(CSRC)/vrt_kernels.o: $(CSRC)/vrt_kernels.cu $(KERNEL_HDRS)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> $(CSRC)/vrt_kernels.ptxas.log || (cat $(CSRC)/vrt_kernels.ptxas.log; false)

$(CSRC)/vrt_shim.o: $(CSRC)/vrt_shim.cu $(KERNEL_HDRS)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<

$(CSRC)/vrt_denoise.o: $(CSRC)/vrt_denoise.cu $(KERNEL_HDRS)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> $(CSRC)/vrt_denoise.ptxas.log || (cat $(CSRC)/vrt_denoise.ptxas.log; false)

$(CSRC)/vrt_build.o: $(CSRC)/vrt_build.cu $(KERNEL_HDRS)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> $(CSRC)/vrt_build.ptxas.log || (cat $(CSRC)/vrt_build.ptxas.log; false)

$(CSRC)/vrt_sched.o: $(CSRC)/vrt_sched.cu $(KERNEL_HDRS)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> $(CSRC)/vrt_sched.ptxas.log || (cat $(CSRC)/vrt_sched.ptxas.log; false)

$(PKG)/libvrt.so: $(CSRC)/vrt_kernels.o $(CSRC)/vrt_shim.o $(CSRC)/vrt_denoise.o $(CSRC)/vrt_build.o $(CSRC)/vrt_sched.o
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $^ -ldl

$(PKG)/libvrt_host.so: $(wildcard $(CSRC)/host/*.cpp) $(wildcard $(CSRC)/host/*.h) include/vrt.h include/vrt_host.h $(PKG)/libvrt.so
	$(CXX) $(HOSTFLAGS) -shared -o $@ $(wildcard $(CSRC)/host/*.cpp) -L$(PKG) -lvrt -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle

# A/B variants of the kernel object for tools/gpu_ab.sh: make ab AB="name1:-DFOO=1 name2:-DBAR=2,-DBAZ=3"
ab: $(PKG)/libvrt.so
	mkdir -p build/ab
	cp $(PKG)/libvrt.so build/ab/libvrt_base.so
	for v in $(AB); do name=$${v%%:*}; defs=$$(echo $${v#*:} | tr ',' ' '); \
	  $(NVCC) $(NVCCFLAGS) $$defs -c -o build/ab/kernels_$$name.o $(CSRC)/vrt_kernels.cu 2> build/ab/kernels_$$name.ptxas.log || exit 1; \
	  $(NVCC) $(NVCCFLAGS) $$defs -c -o build/ab/shim_$$name.o $(CSRC)/vrt_shim.cu 2> /dev/null || exit 1; \
	  $(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/libvrt_$$name.so build/ab/kernels_$$name.o build/ab/shim_$$name.o $(CSRC)/vrt_denoise.o $(CSRC)/vrt_build.o $(CSRC)/vrt_sched.o -ldl || exit 1; \
	  grep -A2 "trace_warp_kernelILi4ELb0ELb1E" build/ab/kernels_$$name.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $$name"; done

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log $(PKG)/*.so
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ab
