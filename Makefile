# Build recipe for the B200-native render path.
#   make lib      libyrt_b200.so (CUDA kernels + C ABI), sm_100a
#   make oracle   oracle/liboracle.so (C restatement, test infrastructure)
#   make ref      oracle/_ref/*  (the UNMODIFIED reference compiled from $(REF)/src where it lies; only in
#                 containers that have $(REF)) and bin/raytrace, bin/yrt_flatten (need the reference's loader)
#   make hostemu  tests/host_emu/libyrt_hostemu.so (device code compiled for the host, tests only)
REF      ?= /root/reference
NVCC     ?= nvcc
CXX      ?= g++
# plain system gcc for the C oracle (the /opt/gcc wrapper some shells export as $CC has no libgomp.spec)
ORACLE_CC ?= $(shell command -v /usr/bin/gcc || echo gcc)
ARCH     := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the parity-critical arithmetic must not be contracted into FMA (SURVEY finding 4)
NVFLAGS  := $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC $(EXTRA)
CSRC     := yocto_raytracing_b200/csrc
HOST     := yocto_raytracing_b200/host
LIB      := yocto_raytracing_b200/libyrt_b200.so
HDRS     := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h include/*.h)
# the reference's own Release flags (build/CMakeCache.txt) + headers modern libstdc++ no longer pulls in
REFFLAGS := -std=c++14 -O3 -DNDEBUG -DYOBJ_NO_IMAGE -DYGLTF_NO_IMAGE -DYSCN_NO_IMAGE -w \
            -include cstring -include stdexcept -include cstdint -include algorithm
REFSRC   := image scene yocto_scn yocto_obj yocto_gltf
REFOBJ   := $(addprefix oracle/_ref/,$(addsuffix .o,$(REFSRC)))

.PHONY: all lib counters oracle ref hostemu clean
all: lib counters oracle hostemu $(if $(wildcard $(REF)/src/raytrace.cpp),ref)

lib: $(LIB)
build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
$(LIB): build/yrt_host.o build/yrt_build.o build/yrt_render.o build/yrt_api.o build/yrt_png.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lz

# the same library with per-ray work counters in the traversal kernels (-DYRT_COUNTERS=1): used by tools/frame_counters.py
# (bench.py runs it in a separate process after its timed region) — never the timed build
COUNTERS_LIB := yocto_raytracing_b200/libyrt_b200_counters.so
counters: $(COUNTERS_LIB)
build/ctr_%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DYRT_COUNTERS=1 -c $< -o $@
$(COUNTERS_LIB): build/ctr_yrt_host.o build/ctr_yrt_build.o build/ctr_yrt_render.o build/ctr_yrt_api.o build/ctr_yrt_png.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lz

oracle: oracle/liboracle.so
oracle/liboracle.so: oracle/yrt_oracle.c oracle/yrt_oracle.h include/yrt_b200.h
	$(ORACLE_CC) -std=c11 -O2 -fPIC -shared -ffp-contract=off -fopenmp -o $@ oracle/yrt_oracle.c -lm

hostemu: tests/host_emu/libyrt_hostemu.so tests/host_emu/libyrt_hostemu_bin.so tests/host_emu/libyrt_hostemu_wide.so
# the same emulation with the other assignments of node arity to ray kind (build options, see yrt_scene.cuh): both ray kinds on
# binary records / both on 4-wide records — keeps every visit routine under test without a GPU
EMUFLAGS := -O2 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off,-fopenmp -shared
tests/host_emu/libyrt_hostemu_bin.so: tests/host_emu/host_emu.cu $(CSRC)/yrt_host.cu $(HDRS)
	$(NVCC) $(EMUFLAGS) -DYRT_WIDE_CLOSEST=2 -DYRT_WIDE_ANY=2 -I$(CSRC) -o $@ tests/host_emu/host_emu.cu $(CSRC)/yrt_host.cu -lgomp
tests/host_emu/libyrt_hostemu_wide.so: tests/host_emu/host_emu.cu $(CSRC)/yrt_host.cu $(HDRS)
	$(NVCC) $(EMUFLAGS) -DYRT_WIDE_CLOSEST=4 -DYRT_WIDE_ANY=4 -I$(CSRC) -o $@ tests/host_emu/host_emu.cu $(CSRC)/yrt_host.cu -lgomp
tests/host_emu/libyrt_hostemu.so: tests/host_emu/host_emu.cu $(CSRC)/yrt_host.cu $(HDRS)
	$(NVCC) $(EMUFLAGS) $(EXTRA) -I$(CSRC) -o $@ tests/host_emu/host_emu.cu $(CSRC)/yrt_host.cu -lgomp

ref: oracle/_ref/raytrace_ref oracle/_ref/ref_probe bin/raytrace bin/yrt_flatten
oracle/_ref/%.o: $(REF)/src/%.cpp
	@mkdir -p oracle/_ref
	$(CXX) $(REFFLAGS) -c $< -o $@
oracle/_ref/%.o: $(REF)/src/ext/%.cpp
	@mkdir -p oracle/_ref
	$(CXX) $(REFFLAGS) -c $< -o $@
oracle/_ref/raytrace_ref: $(REFOBJ) oracle/_ref/raytrace.o
	$(CXX) -o $@ $^
oracle/_ref/ref_probe: oracle/ref_probe.cpp $(REFOBJ)
	$(CXX) $(REFFLAGS) -I$(REF)/src -o $@ $^
bin/raytrace: $(HOST)/raytrace_main.cpp $(HOST)/yrt_flatten.cpp $(HOST)/yrt_flatten.h $(REFOBJ) $(LIB)
	@mkdir -p bin
	$(CXX) $(REFFLAGS) -I$(REF)/src -I$(HOST) -o $@ $(HOST)/raytrace_main.cpp $(HOST)/yrt_flatten.cpp $(REFOBJ) \
	    -L yocto_raytracing_b200 -lyrt_b200 -Wl,-rpath,'$$ORIGIN/../yocto_raytracing_b200'
bin/yrt_flatten: $(HOST)/yrt_flatten_tool.cpp $(HOST)/yrt_flatten.cpp $(HOST)/yrt_flatten.h $(REFOBJ)
	@mkdir -p bin
	$(CXX) $(REFFLAGS) -I$(REF)/src -I$(HOST) -o $@ $(HOST)/yrt_flatten_tool.cpp $(HOST)/yrt_flatten.cpp $(REFOBJ)

clean:
	rm -rf build bin oracle/_ref oracle/liboracle.so tests/host_emu/libyrt_hostemu*.so $(LIB) $(COUNTERS_LIB)
