# Build of the siga_b200 product library, the host CLI, the test-only host stand-in and the oracle.
#   make            -> siga_b200/libsiga_b200.so + siga_b200/bin/siga
#   make hostemu    -> build/libhostemu.so (tests only)
#   make hostcli    -> build/siga_hostemu (tests only: the front-end sources linked against a host stand-in of the ABI)
#   make oracle     -> oracle/liboracle.so (+ oracle/_ref/siga when /root/reference is present)
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   ?= /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 -lineinfo $(ARCH) -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v --expt-relaxed-constexpr
CSRC      := siga_b200/csrc
CU        := $(CSRC)/abi.cu $(CSRC)/reads.cu $(CSRC)/fm.cu $(CSRC)/overlap.cu $(CSRC)/correct.cu $(CSRC)/index.cu $(CSRC)/index_sharded.cu $(CSRC)/comm.cu
OBJ       := $(patsubst $(CSRC)/%.cu,build/obj/%.o,$(CU))
HDR       := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)
LIB       := siga_b200/libsiga_b200.so

all: $(LIB) cli

build/obj/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/obj/$*.ptxas.log || (cat build/obj/$*.ptxas.log; exit 1)
	@grep -E "error|warning" build/obj/$*.ptxas.log | grep -v "ptxas info" || true

$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -ccbin $(HOSTCXX) -o $@.tmp $(OBJ) -lcudart_static -lpthread -ldl -lrt && mv -f $@.tmp $@

cli: siga_b200/bin/siga
siga_b200/bin/siga: $(wildcard $(CSRC)/host/*.cpp $(CSRC)/host/*.h) include/siga_b200.h $(LIB)
	@mkdir -p siga_b200/bin
	@if ls $(CSRC)/host/*.cpp >/dev/null 2>&1; then \
	  $(HOSTCXX) -std=c++17 -O2 -Wall -Wextra -Iinclude $(CSRC)/host/*.cpp -o $@ -Lsiga_b200 -lsiga_b200 -lz -lpthread -ldl -Wl,-rpath,'$$ORIGIN/..'; \
	else echo "no host CLI sources yet"; fi

hostemu: build/libhostemu.so
build/libhostemu.so: tests/hostemu/hostemu.cpp $(HDR)
	@mkdir -p build
	$(HOSTCXX) -std=c++14 -O2 -Wall -Wextra -shared -fPIC $< -o $@

hostcli: build/siga_hostemu
build/siga_hostemu: tests/hostemu/sgx_stub.cpp tests/hostemu/hostemu.cpp $(wildcard $(CSRC)/host/*.cpp $(CSRC)/host/*.h) $(HDR)
	@mkdir -p build
	$(HOSTCXX) -std=c++17 -O2 -Wall -Wextra -Iinclude tests/hostemu/sgx_stub.cpp $(CSRC)/host/*.cpp -o $@ -lz -lpthread -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build/obj $(LIB) siga_b200/bin build/libhostemu.so build/siga_hostemu

.PHONY: all cli hostemu hostcli oracle clean
