"""Developer aid: quick throughput probe of the fused step on a batch of 700^2 environments."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[3]) if len(sys.argv) > 3 else 700
design = int(sys.argv[4]) if len(sys.argv) > 4 else 1
EN = os.environ.get("PERF_ENERGY", "1") == "1"
dim = wb.TwoDim(15.0, n)
eng = wb.Engine(dim.x, dim.y, 1531.0, 1e-5, 2.0, 20000.0, n_env=E)
rng = np.random.default_rng(0)
ds = wb.build_triple_ring_design_space()
shape = wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0])
eng.set_source(shape, 1000.0)
tspan = wb.build_tspan(0.0, 1e-5, steps)
if design:
    for e in range(E):
        d0 = ds.rand(rng)
        d1 = ds(d0, wb.build_action_space(d0, 0.25).rand(rng))
        eng.set_design(d0.table(), d1.table(), tspan[0], tspan[-1], env=e)
u0 = (rng.standard_normal((1, 12, n, n)) * 1e-3).astype(np.float32)
if os.environ.get("PERF_ZERO", "0") == "1":   # start at rest like reset!(env): the auxiliary fields stay zero outside the PML
    u0[:] = 0
for e in range(E):
    eng.set_state(u0, env=e)
for rep in range(2):
    eng.integrate(tspan, wb.MODE_FUSED, energy=EN)
t = time.time()
eng.integrate(tspan, wb.MODE_FUSED, energy=EN)
dt = time.time() - t
cells = E * n * n * steps
print(f"E={E} n={n} steps={steps} design={design}: wall {dt*1e3:.2f} ms  -> {cells/dt/1e9:.2f} Gcell-updates/s  ({cells*96/dt/1e9:.0f} GB/s algorithmic)")
eng.profile(True)
eng.integrate(tspan, wb.MODE_FUSED, energy=EN)
ms, nl = eng.profile_read()
print(f"  per-step kernels (events): {ms/nl*1e3:.1f} us/step -> {E*n*n/(ms/nl*1e-3)/1e9:.2f} Gcell-updates/s, {E*n*n*96/(ms/nl*1e-3)/1e9:.0f} GB/s")
