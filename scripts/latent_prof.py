"""One forward (energies only) and one reverse call of the latent path at 148 x 1024 x 300 steps: the ncu target."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import waves_b200 as wb  # noqa: E402
from latent_cases import make_case  # noqa: E402

steps = int(os.environ.get("LAT_STEPS", 300))
cs = make_case(n=1024, batch=148, steps=steps, nseq=steps // 100 + 1, seed=1)
th = cs["theta"]
it = wb.LatentIntegrator(wb.LatentDynamics(wb.OneDim(cs["dim"].x), 1531.0, 10.0, 10000.0), cs["dt"])
theta = [wb.LinearInterpolation(th.X, th.Y), wb.LatentSource(th.shape, th.freq), th.pml]
it.set_variant(int(os.environ.get("LAT_VARIANT", 0)))   # 0 defaults, 2 pair forward, 4 register reverse, 6 pair reverse
last, e = it(cs["z0"], cs["tspan"], theta, want_z=False, want_energy=True)
print("fwd ms", it.last_kernel_ms())
if os.environ.get("LAT_ONLY_FWD"):
    sys.exit(0)
z = it(cs["z0"], cs["tspan"], theta)
it.adjoint(z, cs["tspan"], theta, w_energy=np.ones((148, 3, steps + 1), np.float32))
print("adj ms", it.last_kernel_ms())
