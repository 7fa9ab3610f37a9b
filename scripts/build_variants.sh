#!/bin/bash
# A/B builds of the warp-queue kernels: one library per parameter set under build/, selected at run time with
# VECCHIO_GPU_LIB.  Both flavours of vk_warpq.cu (trimmed "simple scene" build and general render build) and the render
# build of vk_stepq.cu are compiled with the tag's defines.  usage: scripts/build_variants.sh tag:"-DVKQ_N_FLAT=.. -DVKQ_K_FLAT=.." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -prec-div=false -prec-sqrt=false -DVK_STRICT=0"
C=vecchio_b200/csrc
OTHERS="$C/vk_api.o $C/vk_relayout.o $C/vk_kernels_fast.o $C/vk_kernels_strict.o $C/vk_wavefront_fast.o $C/vk_wavefront_strict.o $C/vk_staged_fast.o $C/vk_staged_strict.o $C/vk_staged_simple.o $C/vk_warpq_strict.o $C/vk_stepq_strict.o $C/vk_kernels_l0.o"
for spec in "$@"; do
  tag="${spec%%:*}"; defs="${spec#*:}"
  ( $NVCC $FLAGS -DVK_SIMPLE=1 $defs -c -o build/wq_$tag.o $C/vk_warpq.cu 2> build/wq_$tag.log || { cat build/wq_$tag.log; exit 1; }
    $NVCC $FLAGS $defs -c -o build/wqf_$tag.o $C/vk_warpq.cu 2> build/wqf_$tag.log || { cat build/wqf_$tag.log; exit 1; }
    $NVCC $FLAGS $defs -c -o build/sq_$tag.o $C/vk_stepq.cu 2> build/sq_$tag.log || { cat build/sq_$tag.log; exit 1; }
    $NVCC $FLAGS -DVK_LIGHT0=1 $defs -c -o build/wql_$tag.o $C/vk_warpq.cu 2> build/wql_$tag.log || { cat build/wql_$tag.log; exit 1; }
    $NVCC $FLAGS -DVK_LIGHT0=1 $defs -c -o build/sql_$tag.o $C/vk_stepq.cu 2> build/sql_$tag.log || { cat build/sql_$tag.log; exit 1; }
    $NVCC $ARCH -shared -o build/libvk_$tag.so $OTHERS build/wq_$tag.o build/wqf_$tag.o build/sq_$tag.o build/wql_$tag.o build/sql_$tag.o
    rm -f build/wq_$tag.o build/wqf_$tag.o build/sq_$tag.o build/wql_$tag.o build/sql_$tag.o # (the gpurun snapshot is capped at 512 MiB; a library is 28 MB)
    echo "$tag: flat-simple $(grep -A2 'k_warpq_flatILb0E' build/wq_$tag.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ') | bvh $(grep -A2 'k_warpqILb0ELb0' build/wqf_$tag.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ') | stepq $(grep -A2 'k_stepqILb0ELb0' build/sq_$tag.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ') inst $(grep -A2 'k_stepq_instILb1ELb0' build/sq_$tag.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')" ) &
done
wait
