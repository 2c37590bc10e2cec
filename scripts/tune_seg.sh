#!/bin/bash
# Runs ON THE GPU BOX: row-slab height (march length) sweep of the fused step at a given batch size (developer build
# build/tune/libwaves_b200_base.so from `scripts/tune_build.sh base ""`).  SEG = clamp(2 ny cols n_env / SEGDIV, 16, SEGCAP).
E=${1:-1024}
shift
for pair in "$@"; do
  cap=${pair%%:*}; div=${pair##*:}
  echo "== SEGCAP $cap SEGDIV $div (E=$E)"
  WAVES_DEBUG_SEGCAP=$cap WAVES_DEBUG_SEGDIV=$div WAVES_B200_LIB=$PWD/build/tune/libwaves_b200_base.so PERF_ZERO=1 timeout 200 python scripts/gpu_perf.py $E 20 2>&1 | tail -1
done
