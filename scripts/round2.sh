#!/bin/bash
# Runs ON THE GPU BOX (gpurun), one call: the whole GPU test suite (no stop on first failure), the bench line with its
# sub-records, the ncu launch list of the same bench command and one ncu --set full capture of a steady-state fused launch set.
# Outputs land in gpurun_out/; scripts/profile_digest.py / scripts/ncu_summary.py turn them into profiles/.
#   gpurun --timeout 1500 -- 'bash scripts/round2.sh r2'
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu_$TAG.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gputest_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest_$TAG.log
tail -5 gpurun_out/gputest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
head -c 1500 gpurun_out/bench_$TAG.json
# the driver's own settings (20 timed steps after 5 warm-up steps: ~20 s of sustained load before the e2e window)
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench20_$TAG.json 2> gpurun_out/bench20_$TAG.err
head -c 600 gpurun_out/bench20_$TAG.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
# launch list of the same command (direct launches: kernel nodes of a replayed graph are listed too)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1
# steady-state launch set (the first step runs the full interior variant that builds the P plane)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_step --launch-skip 8 -c 4 \
    -f -o gpurun_out/prof_$TAG python scripts/gpu_perf.py 128 4 > gpurun_out/ncu_full_$TAG.log 2>&1
PERF_ZERO=1 python scripts/gpu_perf.py 128 20 > gpurun_out/perf128_$TAG.log 2>&1
PERF_ZERO=1 python scripts/gpu_perf.py 1024 20 > gpurun_out/perf1024_$TAG.log 2>&1
cat gpurun_out/perf128_$TAG.log gpurun_out/perf1024_$TAG.log
ls -la gpurun_out | tail -12
