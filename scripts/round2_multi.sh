#!/bin/bash
# Runs ON THE GPU BOX with N GPUs (gpurun --gpus N): the multi-GPU tests (slab decomposition bitwise vs one GPU), the bench line
# at N ranks exactly as the driver launches it (env-sharded batch + the "slab" sub-record: one 16384^2 grid over the N GPUs) and
# the reference arm under torchrun.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/round2_multi.sh 2 r2'
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab_local.py -q -m gpu -p no:cacheprovider "tests/test_gpu_parity.py::test_decimated_trajectory_capture" "tests/test_gpu_parity.py::test_amortised_launches_equal_direct_launches" > gpurun_out/gputest_multi_${TAG}_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest_multi_${TAG}_n$N.log
tail -4 gpurun_out/gputest_multi_${TAG}_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench rc=$?"
cat gpurun_out/bench_${TAG}_n$N.json
tail -3 gpurun_out/bench_${TAG}_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > gpurun_out/bench_ref_${TAG}_n$N.json 2> gpurun_out/bench_ref_${TAG}_n$N.err
cat gpurun_out/bench_ref_${TAG}_n$N.json
