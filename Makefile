# Build recipe for libb200nb (the product), the oracle (test infrastructure) and the reference-side binaries.
# `python -c "import __graft_entry__ as g; g.build()"` runs the same commands.
NVCC      ?= nvcc
CXX       ?= g++
# the image exports CC=/opt/gcc/bin/gcc, which has no libgomp.spec: use the distro compiler for the OpenMP oracle
ORACLE_CC ?= $(shell command -v /usr/bin/gcc || echo gcc)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
PKG       := nbody-eurohpc_b200
LIB       := $(PKG)/b200nb/libb200nb.so
REF       ?= /root/reference

all: lib oracle

# the default force-kernel variant with its hot loop re-ordered after ptxas (tools/sass_resched.py + the committed order)
RESCHED_VARIANT := 32, 8, 2, 2, 1, false, 1, 8, 1
RESCHED_ORDER   := $(PKG)/csrc/resched/pk_t32_r8_tj2_st2_cta_u1_mb8.order.json
RESCHED_INC     := $(PKG)/csrc/generated/force_resched_cubin.inc
$(RESCHED_INC): tools/sass_resched.py $(RESCHED_ORDER) $(PKG)/csrc/force_sm100.cuh $(PKG)/csrc/plan.hpp
	mkdir -p $(dir $@) build/resched && python3 tools/sass_resched.py "$(RESCHED_VARIANT)" build/resched/default.cubin --order-in $(RESCHED_ORDER) --emit-header $@

lib: $(LIB)
$(LIB): $(PKG)/csrc/context.cu $(PKG)/csrc/host_ic.cpp $(PKG)/csrc/force_sm100.cuh $(PKG)/csrc/integrate_sm100.cuh $(PKG)/csrc/plan.hpp include/b200nb.h $(RESCHED_INC)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/context.cu $(PKG)/csrc/host_ic.cpp -ldl

oracle: oracle/liboracle.so
oracle/liboracle.so: oracle/nbody_oracle.c oracle/nbody_oracle.h
	$(ORACLE_CC) -O2 -ffp-contract=off -fopenmp -fPIC -shared -o $@ oracle/nbody_oracle.c -lm || \
	$(ORACLE_CC) -O2 -ffp-contract=off -fPIC -shared -o $@ oracle/nbody_oracle.c -lm

# reference-side artefacts (need $(REF); outputs only under oracle/_ref/)
ref:
	bash oracle/build_ref.sh $(REF)

kbench: build/kbench
build/kbench: tools/kbench.cu $(PKG)/csrc/force_sm100.cuh $(PKG)/csrc/plan.hpp
	mkdir -p build && $(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -o $@ tools/kbench.cu -L/usr/local/cuda/lib64/stubs -lcuda

clean:
	rm -f $(LIB) oracle/liboracle.so build/kbench; rm -rf oracle/_ref

.PHONY: all lib oracle ref kbench clean
