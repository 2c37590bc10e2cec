#!/bin/bash
# ncu --set full of one steady-state step kernel of a single 700^2 environment (direct launches: WAVES graph/coop off)
TAG=${1:-r2}
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_fused_step --launch-skip 30 -c ${2:-1} -o gpurun_out/prof_single_$TAG -f \
    python -c "
import sys, numpy as np
sys.path.insert(0, '.')
import waves_b200 as wb
dim = wb.TwoDim(15.0, 700)
src = wb.RandomPosGaussianSource(dim, [-10.0, -10.0], [-10.0, 10.0], [0.3], [1.0], 1000.0, rng=np.random.default_rng(1))
env = wb.WaveEnv(dim, design_space=wb.build_triple_ring_design_space(), source=src, integration_steps=100, actions=3, rng=np.random.default_rng(2))
env.iter.engine.set_graph(False)
env(env.action_space().rand(np.random.default_rng(0)))
" > gpurun_out/ncu_single_$TAG.log 2>&1
tail -2 gpurun_out/ncu_single_$TAG.log
