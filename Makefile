# Builds libce2e.so (the C ABI of include/ce2e.h) and the plain-C client without Python.
# Same flags as env_build_b200/_lib.py (tests/test_host_cpu.py::test_makefile_flags_match keeps them in step).
NVCC      ?= nvcc
CC        ?= gcc
CUDA_HOME ?= /usr/local/cuda
CSRC      := env_build_b200/csrc
LIB       := $(CSRC)/libce2e.so
NVCC_FLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC,-ffp-contract=off

all: $(LIB)

$(LIB): $(CSRC)/ce2e.cu $(CSRC)/ce2e_device.cuh $(CSRC)/ce2e_grid.h include/ce2e.h
	$(NVCC) $(NVCC_FLAGS) -I include -I $(CSRC) -o $@ $<

# tests/c_abi/kat.c: known answers through the C ABI (needs a GPU to run)
build/kat: tests/c_abi/kat.c include/ce2e.h $(LIB)
	mkdir -p build
	$(CC) -std=c99 -Wall -I include -isystem $(CUDA_HOME)/include $< -o $@ -L $(CSRC) -l:libce2e.so \
	    -L $(CUDA_HOME)/lib64 -lcudart -lm -Wl,-rpath,$(abspath $(CSRC)) -Wl,-rpath,$(CUDA_HOME)/lib64

c_client: build/kat

clean:
	rm -f $(LIB) build/kat

.PHONY: all c_client clean
