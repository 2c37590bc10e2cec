#!/bin/bash
# Runs ON THE GPU BOX (gpurun), one call, about 2 GPU-minutes: everything the 1-D latent row still needs from
# a B200 -- the GPU tests of the opt-in kernels, timings of every variant next to the defaults, and one sectioned ncu
# capture per opt-in kernel.  Outputs land in gpurun_out/ (copy the summaries you keep into profiles/).
#   gpurun --timeout 240 -- 'bash scripts/latent_round.sh r2'
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_zlatent.py -q > gpurun_out/latent_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/latent_tests_$TAG.log
tail -3 gpurun_out/latent_tests_$TAG.log
timeout 60 python scripts/bench_latent.py > gpurun_out/latent_bench_$TAG.jsonl 2> gpurun_out/latent_bench_$TAG.err
cat gpurun_out/latent_bench_$TAG.jsonl
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats"
# forward pair kernel (energies only) and both register reverse kernels; -c 1 each: ncu replays a kernel ~40 times
LAT_VARIANT=2 LAT_ONLY_FWD=1 timeout 60 ncu $SECTIONS --clock-control none --import-source on -k regex:k_latent_integrate_r2 -c 1 \
    -o gpurun_out/prof_latent_fwd_pair_$TAG -f python scripts/latent_prof.py > gpurun_out/ncu_latent_fwd_pair_$TAG.log 2>&1
LAT_VARIANT=4 timeout 60 ncu $SECTIONS --clock-control none --import-source on -k regex:k_latent_adjoint_r1 -c 1 \
    -o gpurun_out/prof_latent_adj_r1_$TAG -f python scripts/latent_prof.py > gpurun_out/ncu_latent_adj_r1_$TAG.log 2>&1
LAT_VARIANT=6 timeout 60 ncu $SECTIONS --clock-control none --import-source on -k regex:k_latent_adjoint_r2 -c 1 \
    -o gpurun_out/prof_latent_adj_pair_$TAG -f python scripts/latent_prof.py > gpurun_out/ncu_latent_adj_pair_$TAG.log 2>&1
ls -la gpurun_out | tail -8
