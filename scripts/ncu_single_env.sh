#!/bin/bash
# per-launch durations of ONE env(action) of a single 700^2 environment (kernels launched directly, no graph)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${1:-r2}_single_env_launches.csv \
    python -c "
import sys, numpy as np
sys.path.insert(0, '.')
import waves_b200 as wb
dim = wb.TwoDim(15.0, 700)
src = wb.RandomPosGaussianSource(dim, [-10.0, -10.0], [-10.0, 10.0], [0.3], [1.0], 1000.0, rng=np.random.default_rng(1))
env = wb.WaveEnv(dim, design_space=wb.build_triple_ring_design_space(), source=src, integration_steps=100, actions=3, rng=np.random.default_rng(2))
env.iter.engine.set_graph(False)
rng = np.random.default_rng(0)
env(env.action_space().rand(rng))
" > gpurun_out/${1:-r2}_single_env_ncu.log 2>&1
python - <<'PY'
import csv, collections, sys
rows = [r for r in csv.reader(open('gpurun_out/${1:-r2}_single_env_launches.csv')) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]
kn, mv = H.index('Kernel Name'), H.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    try:
        agg[r[kn][:60]].append(float(r[mv].replace(',', '')))
    except Exception:
        pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us total={sum(v)/1e6:8.3f} ms")
PY
