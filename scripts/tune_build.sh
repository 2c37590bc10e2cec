#!/bin/bash
# Developer aid: build tuning variants of libwaves_b200.so into build/tune/ (git-ignored; they travel with gpurun).
# These are -DWAVES_DEV builds: they honour the WAVES_DEBUG_* environment switches and report WAVES_BUILD_DEV from
# waves_build_flags(), so bench.py refuses them (select one with WAVES_B200_LIB for scripts/gpu_perf.py and friends).
#   scripts/tune_build.sh tag "-DWV_P_REGS=0 -DWV_OCC_STRIP=12" ...
set -e
cd "$(dirname "$0")/../waves.jl_b200/csrc"
OUT=../../build/tune
mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off -I../../include -DWAVES_DEV"
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  $NV $flags -Xptxas -v -c -o $OUT/fused_$tag.o kernels_fused.cu 2> $OUT/fused_$tag.ptxas.log
  $NV -c -o $OUT/abi_dev.o waves_abi.cu
  $NV -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libwaves_b200_$tag.so kernels_exact.o kernels_adjoint.o kernels_adjoint_fused.o $OUT/fused_$tag.o $OUT/abi_dev.o latent_abi.o -lcudart_static -lpthread -ldl -lrt
  echo "$tag: $(grep -E 'Used [0-9]+ registers' $OUT/fused_$tag.ptxas.log | head -4 | sed 's/ptxas info    : Used //; s/ registers.*//' | tr '\n' ' ')"
done
