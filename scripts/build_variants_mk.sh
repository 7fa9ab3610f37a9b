#!/bin/bash
# A/B builds of the lane megakernels (vk_kernels.cu, render flavours vkfast and vkfast_l0): one library per parameter set under
# build/, selected at run time with VECCHIO_GPU_LIB.  usage: scripts/build_variants_mk.sh tag:"-DVK_MINB_BVH=8" tag2:"-Xptxas -O2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -prec-div=false -prec-sqrt=false -DVK_STRICT=0 -DVKD_MEDIUM_SPAN=0"
C=vecchio_b200/csrc
OTHERS="$C/vk_api.o $C/vk_relayout.o $C/vk_kernels_strict.o $C/vk_wavefront_fast.o $C/vk_wavefront_strict.o $C/vk_staged_fast.o $C/vk_staged_strict.o $C/vk_staged_simple.o $C/vk_warpq_fast.o $C/vk_warpq_strict.o $C/vk_warpq_simple.o $C/vk_stepq_fast.o $C/vk_stepq_strict.o $C/vk_warpq_l0.o $C/vk_stepq_l0.o"
for spec in "$@"; do
  tag="${spec%%:*}"; defs="${spec#*:}"
  ( $NVCC $FLAGS $defs -c -o build/mk_$tag.o $C/vk_kernels.cu 2> build/mk_$tag.log || { cat build/mk_$tag.log; exit 1; }
    $NVCC $FLAGS -DVK_LIGHT0=1 $defs -c -o build/mkl_$tag.o $C/vk_kernels.cu 2> build/mkl_$tag.log || { cat build/mkl_$tag.log; exit 1; }
    $NVCC $ARCH -shared -o build/libvk_$tag.so $OTHERS build/mk_$tag.o build/mkl_$tag.o
    rm -f build/mk_$tag.o build/mkl_$tag.o
    echo "$tag: l0 k_megakernel<1,0> $(grep -A2 'k_megakernelILb1ELb0' build/mkl_$tag.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ') | <0,0> $(grep -A2 'k_megakernelILb0ELb0' build/mkl_$tag.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')" ) &
done
wait
