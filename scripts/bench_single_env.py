"""BASELINE configs[1] latency: ONE 700^2 WaveEnv (triple-ring design + Gaussian source), env(action) = 100 RK4 steps + energy
trace, in the three launch modes of waves_set_graph (default: one launch per step for a small batch).  Prints one JSON line.
  python scripts/bench_single_env.py [actions]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb  # noqa: E402


def measure(actions=20, steps=100, n=700, device=0):
    dim = wb.TwoDim(15.0, n)
    rng = np.random.default_rng(0)
    out = {}
    for name, graph in (("default", 1), ("cooperative_multi_step", 2), ("direct", 0)):
        src = wb.RandomPosGaussianSource(dim, [-10.0, -10.0], [-10.0, 10.0], [0.3], [1.0], 1000.0, rng=np.random.default_rng(1))
        env = wb.WaveEnv(dim, design_space=wb.build_triple_ring_design_space(), source=src, integration_steps=steps,
                         actions=actions + 3, rng=np.random.default_rng(2), device=device)
        env.iter.engine.set_graph(graph)
        for _ in range(3):
            env(env.action_space().rand(rng))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(actions):
            env(env.action_space().rand(rng))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"us_per_rk4_step": round(dt / (actions * steps) * 1e6, 2), "ms_per_action": round(dt / actions * 1e3, 3),
                     "launches_per_action": env.iter.engine.launch_count() // (actions + 3)}
        env.iter.engine.close()
    return out


if __name__ == "__main__":
    a = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    print(json.dumps({"workload": "one 700^2 WaveEnv, env(action) = 100 fused RK4 steps + 101x3 energy trace, host wall clock per action",
                      **measure(a)}))
