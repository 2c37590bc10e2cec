"""Timing of the 1-D latent path on one B200: kernel time from CUDA events inside the library
(`waves_latent_last_kernel_ms`), host-pointer call time by wall clock.  Prints one JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import waves_b200 as wb  # noqa: E402
from latent_cases import make_case  # noqa: E402

F32 = np.float32
if "--cpu" in sys.argv:
    # the oracle (NumPy restatement of the reference's array program, one thread) on the host cores: a bounded sample
    from oracle import latent_oracle as lo
    cs = make_case(n=1024, batch=32, steps=100, nseq=2, seed=1)
    t = time.perf_counter()
    z = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    lo.compute_latent_energy(z, 1.0)
    dtm = time.perf_counter() - t
    print(json.dumps({"cpu_oracle": {"batch": 32, "steps": 100, "n": 1024, "seconds": round(dtm, 3),
                                     "us_per_step": round(1e6 * dtm / 100, 1),
                                     "Melement_steps_per_s": round(32 * 1024 * 100 / dtm / 1e6, 1), "cores": 1,
                                     "kind": "port (oracle/latent_oracle.py, NumPy float32)"}}))
    sys.exit(0)
out = []
for batch, steps in ((32, 100), (148, 300), (1184, 300), (148, 2000)):
    cs = make_case(n=1024, batch=batch, steps=steps, nseq=steps // 100 + 1, seed=1)
    th = cs["theta"]
    it = wb.LatentIntegrator(wb.LatentDynamics(wb.OneDim(cs["dim"].x), 1531.0, 10.0, 10000.0), cs["dt"])
    theta = [wb.LinearInterpolation(th.X, th.Y), wb.LatentSource(th.shape, th.freq), th.pml]
    rec = {"batch": batch, "steps": steps, "n": 1024}
    for name, kw in (("energy_only", dict(want_z=False, want_energy=True)), ("trajectory", dict(want_z=True, want_energy=True)),
                     ("energy_only_pair_variant", dict(want_z=False, want_energy=True))):
        if name == "trajectory" and batch * steps > 148 * 300:
            continue
        it.set_variant(wb.LATENT_PAIR if name.endswith("pair_variant") else wb.LATENT_SINGLE)
        ms = []
        for _ in range(4):
            t = time.perf_counter()
            res = it(cs["z0"], cs["tspan"], theta, **kw)
            wall = (time.perf_counter() - t) * 1e3
            ms.append((it.last_kernel_ms(), wall))
        k, w = min(m[0] for m in ms[1:]), min(m[1] for m in ms[1:])
        rec[name] = {"kernel_ms": round(k, 4), "call_ms_host_buffers": round(w, 3), "us_per_step": round(1e3 * k / steps, 3),
                     "Melement_steps_per_s": round(batch * 1024 * steps / k / 1e3, 1)}
    it.set_variant(wb.LATENT_AUTO)
    if batch * steps <= 148 * 300:
        z = it(cs["z0"], cs["tspan"], theta)
        wE = np.ones((batch, 3, steps + 1), F32)
        for name, variant in (("adjoint", wb.LATENT_SINGLE), ("adjoint_register_variant", wb.LATENT_SINGLE | wb.LATENT_ADJ_R1),
                              ("adjoint_register_pair_variant", wb.LATENT_ADJ_R1 | wb.LATENT_PAIR)):
            it.set_variant(variant)
            for _ in range(3):
                it.adjoint(z, cs["tspan"], theta, w_energy=wE)
            rec[name] = {"kernel_ms": round(it.last_kernel_ms(), 4), "us_per_step": round(1e3 * it.last_kernel_ms() / steps, 3)}
        it.set_variant(wb.LATENT_AUTO)
    out.append(rec)
    print(json.dumps(rec), flush=True)
    it.close()
