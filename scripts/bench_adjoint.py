"""BASELINE configs[4]: gradient of the scattered energy through N steps (forward + reverse RK4 kernels) on one 700^2
environment with the triple-ring design frozen.  `python scripts/bench_adjoint.py [steps]`"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # environments differentiated at once (each stores steps+1 states)
n = 700
dim = wb.TwoDim(15.0, n)
eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=E)
eng.set_source(wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0]), 1000.0)
rng = np.random.default_rng(0)
d0 = wb.build_triple_ring_design_space().rand(rng)
ts = wb.build_tspan(0.0, 1e-5, steps)
eng.set_design(d0.table(), d0.table(), ts[0], ts[-1])
w = np.zeros((steps + 1, 3), np.float32)
w[:, 2] = 1.0   # L = sum_t E_sc(t)
z0 = np.zeros((E, 12, n, n), np.float32)
gz_d = torch.empty((E, 12, n, n), dtype=torch.float32, device="cuda")   # gradients stay on the device
gc_d = torch.empty((E, n, n), dtype=torch.float32, device="cuda")
res = {}
for mode, name in ((wb.ADJ_EXACT, "exact"), (wb.ADJ_COMPAT, "compat")):
    for rep in range(2):
        eng.set_state(z0)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        loss, gz, gc = eng.adjoint(ts, w, adj_mode=mode, out_dz0=gz_d, out_dc=gc_d)
        dt = time.perf_counter() - t0
    res[name] = {"seconds": round(dt, 3), "Gcell_updates_per_s_fwd_plus_rev": round(2 * E * n * n * steps / dt / 1e9, 3),
                 "launches": eng.launch_count() - l0, "loss": float(loss[0]), "norm_dL_dc": float(gc.norm().item()),
                 "finite": bool(torch.isfinite(gc).all().item() and torch.isfinite(gz).all().item())}
# dL/dz0 only (the reference-equivalent gradient): no forward-stage recomputation, 4 launches per reverse step
for rep in range(2):
    eng.set_state(z0)
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    loss, gz, _ = eng.adjoint(ts, w, adj_mode=wb.ADJ_EXACT, out_dz0=gz_d, want_dc=False)
    dt = time.perf_counter() - t0
res["exact_z0_only"] = {"seconds": round(dt, 3), "Gcell_updates_per_s_fwd_plus_rev": round(2 * E * n * n * steps / dt / 1e9, 3),
                        "launches": eng.launch_count() - l0, "finite": bool(torch.isfinite(gz).all().item())}
print(json.dumps({"workload": f"{E} x 700^2, triple-ring design frozen, {steps} steps, L = sum_t E_sc(t): forward (fused) + reverse sweep",
                  **res}))
eng.close()
