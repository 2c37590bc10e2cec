#!/bin/bash
# Runs ON THE GPU BOX: row-slab height (march length) sweep of the fused step at a given batch size (developer build)
E=${1:-1024}
for cap in 192 300 400 600; do
  echo "== SEGCAP $cap (E=$E)"
  WAVES_DEBUG_SEGCAP=$cap WAVES_B200_LIB=$PWD/build/tune/libwaves_b200_base.so PERF_ZERO=1 timeout 200 python scripts/gpu_perf.py $E 10 2>&1 | tail -2
done
