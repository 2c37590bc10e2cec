#!/bin/bash
# ncu --set full of the two fused reverse-step kernels (32 x 700^2, steady state)
TAG=${1:-r2}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_adjoint_step --launch-skip 6 -c 2 -o gpurun_out/prof_adj_$TAG -f \
    python -c "
import sys; sys.path.insert(0, 'scripts'); import bench_adjoint
print(bench_adjoint.measure(12, 32, with_dc=False, check_exact=False))" > gpurun_out/ncu_adj_$TAG.log 2>&1
tail -3 gpurun_out/ncu_adj_$TAG.log
