# Build of the three shared libraries and the frame-loop program.  `make` = all; `make gpu` needs nvcc only (no GPU).
#   vecchio_b200/lib/libvecchio_gpu.so   CUDA kernels + C ABI (include/vecchio_gpu.h)   -- product
#   vecchio_b200/lib/libvecchio_host.so  host front end (reference scene API + lower())  -- product
#   vecchio_b200/lib/vecchio_gpu_render  the reference's main() over the C ABI (frame loop, P3 output) -- product
#   oracle/liboracle.so                  CPU restatement of the reference                -- tests only
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
TUNE      ?=
WQ_SIMPLE ?=
NVFLAGS   := $(TUNE) $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS  := -std=c++17 -O3 -fPIC -ffp-contract=off -Wall -Wextra
LIBDIR    := vecchio_b200/lib
CSRC      := vecchio_b200/csrc
HOSTSRC   := vecchio_b200/host

all: host oracle gpu render

host: $(LIBDIR)/libvecchio_host.so
oracle: oracle/liboracle.so
gpu: $(LIBDIR)/libvecchio_gpu.so
render: $(LIBDIR)/vecchio_gpu_render

$(LIBDIR)/libvecchio_host.so: $(HOSTSRC)/vecchio.cpp $(HOSTSRC)/scene.cpp $(HOSTSRC)/capi.cpp $(HOSTSRC)/vecchio.hpp include/vecchio_gpu.h include/vecchio_host.h
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOSTSRC)/vecchio.cpp $(HOSTSRC)/scene.cpp $(HOSTSRC)/capi.cpp -lz

oracle/liboracle.so: oracle/oracle.cpp oracle/oracle.h include/vecchio_gpu.h
	$(CXX) $(CXXFLAGS) -fopenmp -shared -o $@ oracle/oracle.cpp

KDEPS := $(CSRC)/vk_device.cuh $(CSRC)/vk_internal.h $(CSRC)/vk_relayout.h include/vecchio_gpu.h Makefile

# the device code is compiled twice: contracted FMA ("fast") and -fmad=false ("strict", the
# reference's op sequence, used for hit parity)
# (the lane megakernels keep medium_t's two boundary queries as two calls: the one-evaluation form of the queue kernels
# is 13 % faster on Cornell smoke there and 2.3 % SLOWER here -- final scene 47.2 against 46.1 ms, profiles/r2_sweep_16.log:
# less arithmetic, but a larger body in a kernel whose top stall is instruction fetch)
$(CSRC)/vk_kernels_fast.o: $(CSRC)/vk_kernels.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVKD_MEDIUM_SPAN=0 -c -o $@ $< 2> $(CSRC)/ptxas_fast.log || (cat $(CSRC)/ptxas_fast.log; false)
# the render build once more for scenes whose light list is one unflipped Rect (VK_LIGHT0, namespace vkfast_l0; see vk_device.cuh)
$(CSRC)/vk_kernels_l0.o: $(CSRC)/vk_kernels.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVKD_MEDIUM_SPAN=0 -DVK_LIGHT0=1 -c -o $@ $< 2> $(CSRC)/ptxas_l0.log || (cat $(CSRC)/ptxas_l0.log; false)
$(CSRC)/vk_kernels_strict.o: $(CSRC)/vk_kernels.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -fmad=false -DVK_STRICT=1 -c -o $@ $< 2> $(CSRC)/ptxas_strict.log || (cat $(CSRC)/ptxas_strict.log; false)
$(CSRC)/vk_wavefront_fast.o: $(CSRC)/vk_wavefront.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -c -o $@ $< 2> $(CSRC)/ptxas_wf_fast.log || (cat $(CSRC)/ptxas_wf_fast.log; false)
$(CSRC)/vk_wavefront_strict.o: $(CSRC)/vk_wavefront.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -fmad=false -DVK_STRICT=1 -c -o $@ $< 2> $(CSRC)/ptxas_wf_strict.log || (cat $(CSRC)/ptxas_wf_strict.log; false)
$(CSRC)/vk_staged_fast.o: $(CSRC)/vk_staged.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -c -o $@ $< 2> $(CSRC)/ptxas_staged_fast.log || (cat $(CSRC)/ptxas_staged_fast.log; false)
# the trimmed build needs fewer registers: 768 slots (55 KB) x 4 CTAs per SM, 3 rays per thread in extend
# (measured: Cornell 49.6 -> 48.4 ms, Cornell smoke 39.4 -> 35.9 ms at 500 spp)
$(CSRC)/vk_staged_simple.o: $(CSRC)/vk_staged.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVK_SIMPLE=1 -DVKS_N=768 -DVKS_K=3 -DVKS_MINB=4 -c -o $@ $< 2> $(CSRC)/ptxas_staged_simple.log || (cat $(CSRC)/ptxas_staged_simple.log; false)
$(CSRC)/vk_staged_strict.o: $(CSRC)/vk_staged.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -fmad=false -DVK_STRICT=1 -c -o $@ $< 2> $(CSRC)/ptxas_staged_strict.log || (cat $(CSRC)/ptxas_staged_strict.log; false)
# the warp-queue kernels (slots, rays per lane and CTAs per SM are set per kernel family in vk_warpq.cu)
KDEPS_WQ := $(KDEPS) $(CSRC)/vk_warpq.cuh
# the step-queue kernels for BVH scenes
$(CSRC)/vk_stepq_fast.o: $(CSRC)/vk_stepq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -c -o $@ $< 2> $(CSRC)/ptxas_stepq_fast.log || (cat $(CSRC)/ptxas_stepq_fast.log; false)
$(CSRC)/vk_stepq_l0.o: $(CSRC)/vk_stepq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVK_LIGHT0=1 -c -o $@ $< 2> $(CSRC)/ptxas_stepq_l0.log || (cat $(CSRC)/ptxas_stepq_l0.log; false)
$(CSRC)/vk_warpq_l0.o: $(CSRC)/vk_warpq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVK_LIGHT0=1 -c -o $@ $< 2> $(CSRC)/ptxas_warpq_l0.log || (cat $(CSRC)/ptxas_warpq_l0.log; false)
$(CSRC)/vk_stepq_strict.o: $(CSRC)/vk_stepq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -fmad=false -DVK_STRICT=1 -c -o $@ $< 2> $(CSRC)/ptxas_stepq_strict.log || (cat $(CSRC)/ptxas_stepq_strict.log; false)
$(CSRC)/vk_warpq_fast.o: $(CSRC)/vk_warpq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -c -o $@ $< 2> $(CSRC)/ptxas_warpq_fast.log || (cat $(CSRC)/ptxas_warpq_fast.log; false)
$(CSRC)/vk_warpq_simple.o: $(CSRC)/vk_warpq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVK_SIMPLE=1 $(WQ_SIMPLE) -c -o $@ $< 2> $(CSRC)/ptxas_warpq_simple.log || (cat $(CSRC)/ptxas_warpq_simple.log; false)
$(CSRC)/vk_warpq_strict.o: $(CSRC)/vk_warpq.cu $(KDEPS_WQ)
	$(NVCC) $(NVFLAGS) -fmad=false -DVK_STRICT=1 -c -o $@ $< 2> $(CSRC)/ptxas_warpq_strict.log || (cat $(CSRC)/ptxas_warpq_strict.log; false)
$(CSRC)/vk_api.o: $(CSRC)/vk_api.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(CSRC)/vk_relayout.o: $(CSRC)/vk_relayout.cu $(KDEPS)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIBDIR)/libvecchio_gpu.so: $(CSRC)/vk_api.o $(CSRC)/vk_relayout.o $(CSRC)/vk_kernels_fast.o $(CSRC)/vk_kernels_strict.o $(CSRC)/vk_wavefront_fast.o $(CSRC)/vk_wavefront_strict.o $(CSRC)/vk_staged_fast.o $(CSRC)/vk_staged_strict.o $(CSRC)/vk_staged_simple.o $(CSRC)/vk_warpq_fast.o $(CSRC)/vk_warpq_strict.o $(CSRC)/vk_warpq_simple.o $(CSRC)/vk_stepq_fast.o $(CSRC)/vk_stepq_strict.o $(CSRC)/vk_kernels_l0.o $(CSRC)/vk_warpq_l0.o $(CSRC)/vk_stepq_l0.o
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^

# the program finds its two libraries next to itself
$(LIBDIR)/vecchio_gpu_render: $(HOSTSRC)/main.cpp include/vecchio_gpu.h include/vecchio_host.h $(LIBDIR)/libvecchio_host.so $(LIBDIR)/libvecchio_gpu.so
	$(CXX) $(CXXFLAGS) -fPIE -o $@ $(HOSTSRC)/main.cpp -L$(LIBDIR) -lvecchio_host -lvecchio_gpu -pthread -Wl,-rpath,'$$ORIGIN'

clean:
	rm -f $(LIBDIR)/*.so $(LIBDIR)/vecchio_gpu_render oracle/*.so $(CSRC)/*.o $(CSRC)/*.log

.PHONY: all host oracle gpu render clean
