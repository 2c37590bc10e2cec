#!/bin/bash
# Runs ON THE GPU BOX (gpurun): bench line, ncu launch list of the same command, one ncu --set full capture of the fused
# launch set.  Outputs land in gpurun_out/; scripts/profile_digest.py turns them into profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
tail -c 600 gpurun_out/bench_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
# skip the first step (it runs the full interior variant that builds the P plane): capture the steady-state launch set
WAVES_DEBUG_FLAGS=16 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_step --launch-skip 4 -c 4 \
    -o gpurun_out/prof_$TAG python scripts/gpu_perf.py 128 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -8
