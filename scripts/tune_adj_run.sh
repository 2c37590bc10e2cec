#!/bin/bash
# Runs ON THE GPU BOX: fused-reverse tests + scripts/bench_adjoint.py (32 x 700^2 x 100 steps) once per tuning build given by tag.
mkdir -p gpurun_out
for tag in "$@"; do
  echo "== $tag"
  export WAVES_B200_LIB=$PWD/build/tune/libwaves_b200_$tag.so
  timeout 300 python -m pytest tests/test_gpu_adjoint_fused.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -1
  timeout 120 python -c "
import sys, json; sys.path.insert(0, 'scripts'); import bench_adjoint
r = bench_adjoint.measure(100, 32, with_dc=False, check_exact=False)
print(json.dumps({k: r[k] for k in ('forward_only_seconds', 'seconds', 'value', 'reverse_Gcell_per_s')}))" 2>&1 | tail -1
done
unset WAVES_B200_LIB
