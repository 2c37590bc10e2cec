#!/bin/bash
# Runs ON THE GPU BOX: scripts/gpu_perf.py (128 x 700^2, 20 steps) once per tuning build in build/tune/ given by tag.
mkdir -p gpurun_out
for tag in "$@"; do
  echo "== $tag"
  WAVES_B200_LIB=$PWD/build/tune/libwaves_b200_$tag.so PERF_ZERO=1 timeout 120 python scripts/gpu_perf.py 128 20 2>&1 | tail -2
done
