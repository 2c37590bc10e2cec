#!/bin/bash
# Runs ON THE GPU BOX: parity tests of the fused path, then the throughput probes (128 / 1024 environments, single environment)
# with the release library and with the tuning builds named on the command line.
#   gpurun --timeout 900 -- 'bash scripts/perf_round.sh TAG [tune tags...]'
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab_local.py tests/test_gpu_adjoint_fused.py -q -x -p no:cacheprovider > gpurun_out/parity_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/parity_$TAG.log; tail -3 gpurun_out/parity_$TAG.log
{
PERF_ZERO=1 python scripts/gpu_perf.py 128 20
PERF_ZERO=1 python scripts/gpu_perf.py 1024 20
python scripts/bench_single_env.py 10
for tag in "$@"; do
  echo "== $tag"
  WAVES_B200_LIB=$PWD/build/tune/libwaves_b200_$tag.so PERF_ZERO=1 timeout 120 python scripts/gpu_perf.py 128 20 2>&1 | tail -2
  WAVES_B200_LIB=$PWD/build/tune/libwaves_b200_$tag.so PERF_ZERO=1 timeout 120 python scripts/gpu_perf.py 1024 20 2>&1 | tail -2
done
} > gpurun_out/perf_$TAG.log 2>&1
cat gpurun_out/perf_$TAG.log
echo "== no design (upper bound of the in-kernel speed field cost)" >> gpurun_out/perf_$TAG.log
PERF_ZERO=1 python scripts/gpu_perf.py 128 20 700 0 >> gpurun_out/perf_$TAG.log 2>&1
tail -3 gpurun_out/perf_$TAG.log
