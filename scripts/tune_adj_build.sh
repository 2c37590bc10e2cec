#!/bin/bash
# Developer aid: tuning variants of the fused reverse-step kernels into build/tune/ (see scripts/tune_build.sh).
#   scripts/tune_adj_build.sh tag "-DWV_ADJ_TX_INT=32 -DWV_ADJ_NT_INT=256" ...
set -e
cd "$(dirname "$0")/../waves.jl_b200/csrc"
OUT=../../build/tune
mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off -I../../include -DWAVES_DEV"
$NV -c -o $OUT/abi_dev.o waves_abi.cu
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  $NV $flags -Xptxas -v -c -o $OUT/adjf_$tag.o kernels_adjoint_fused.cu 2> $OUT/adjf_$tag.ptxas.log
  $NV -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libwaves_b200_$tag.so kernels_exact.o kernels_adjoint.o kernels_fused.o $OUT/adjf_$tag.o $OUT/abi_dev.o latent_abi.o -lcudart_static -lpthread -ldl -lrt
  echo "$tag: $(grep -E 'Used [0-9]+ registers|spill' $OUT/adjf_$tag.ptxas.log | tr '\n' ' ' | sed 's/ptxas info    ://g')"
done
