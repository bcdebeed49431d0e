#!/bin/bash
# build_variant.sh NAME [extra nvcc flags for imhd_fused.cu ...]  ->  tools/experiments/_build/libimhd_NAME.so
# (the library with another build of the fused kernels; select it with IMHD_B200_LIB=... for tools/ab_bench.py)
set -e
cd "$(dirname "$0")/../../imhd-cuda_b200/csrc"
name=$1; shift
src=${IMHD_FUSED_SRC:-imhd_fused.cu}
out=../../tools/experiments/_build
mkdir -p $out
make -s > /dev/null 2>&1
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" -c $src -o $out/fused_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libimhd_$name.so imhd_api.o imhd_granular.o $out/fused_$name.o imhd_stability.o imhd_slabs.o imhd_h5.o -lcudart -ldl
rm -f $out/fused_$name.o
echo built $out/libimhd_$name.so
