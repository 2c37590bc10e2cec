"""Randomized sweep of the latent kernels under the host emulation (tests/emu/): every forward kernel must equal the oracle
bit for bit, every reverse kernel must equal the generic one to float32 rounding.  Developer aid (CPU only)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_latent_cpu as T  # noqa: E402
from latent_cases import make_case  # noqa: E402
from oracle import latent_oracle as lo  # noqa: E402
from oracle import waves_oracle as wo  # noqa: E402

F32 = np.float32
L = C.CDLL(os.path.join(ROOT, "tests", "emu", "liblatent_emu.so"))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = 0
for case in range(ncase):
    knots = ["actions", "partial", "repeated"][rng.integers(3)]
    nseq = 4 if knots == "repeated" else int(rng.integers(2, 6))
    steps = int((nseq - 1) * rng.integers(2, 7)) if knots != "partial" else int(rng.integers(6, 20))
    if knots == "repeated":
        steps = int(2 * rng.integers(3, 8))
    n = int(rng.integers(3, 260))
    batch = int(rng.integers(1, 3))
    cs = make_case(n=n, batch=batch, steps=steps, nseq=nseq, seed=int(rng.integers(1 << 30)), knots=knots,
                   t0=float(rng.uniform(0, 5e-3)))
    nt1 = ((n + 31) // 32) * 32
    want = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    we = lo.compute_latent_energy(want, wo.get_dx(cs["dim"]))
    kernels = [("generic", L.emu_latent_integrate, [nt1, 32, 64][rng.integers(3)]), ("r1", L.emu_latent_integrate_r1, nt1)]
    if n % 2 == 0 and n >= 4:
        kernels.append(("r2", L.emu_latent_integrate_r2, ((n // 2 + 31) // 32) * 32))
    for name, fn, nt in kernels:
        z = np.full(want.shape, np.nan, F32)
        e = np.full(we.shape, np.nan, F32)
        p, keep = T._params(cs, z=z, energy=e)
        fn(C.byref(p), int(nt))
        ok = np.array_equal(z, want) and np.allclose(e, we, rtol=2e-6, atol=2e-6 * we.max())
        if not ok:
            bad += 1
            print("FORWARD MISMATCH", name, dict(n=n, batch=batch, steps=steps, nseq=nseq, knots=knots, nt=int(nt)))
    wE = rng.standard_normal((batch, 3, steps + 1)).astype(F32)
    dz = (1e-2 * rng.standard_normal(want.shape)).astype(F32) if rng.integers(2) else None
    res = {}
    akern = [("generic", L.emu_latent_adjoint, nt1), ("r1", L.emu_latent_adjoint_r1, nt1)]
    if n % 2 == 0 and n >= 4:
        akern.append(("r2", L.emu_latent_adjoint_r2, ((n // 2 + 31) // 32) * 32))
    compat = int(rng.integers(2))
    for name, fn, nt in akern:
        g = dict(z0=np.full((batch, 4, n), np.nan, F32), Y=np.zeros((batch, nseq, n), F32), shape=np.full((batch, n), np.nan, F32),
                 pml=np.full((batch, n), np.nan, F32))
        p, keep = T._params(cs, zt=want, w_energy=wE, dL_dz=dz, g_z0=g["z0"], g_Y=g["Y"], g_shape=g["shape"], g_pml=g["pml"])
        p.compat = compat
        fn(C.byref(p), int(nt))
        res[name] = g
    for name in res:
        if name == "generic":
            continue
        for k in ("z0", "Y", "shape", "pml"):
            den = np.linalg.norm(res["generic"][k])
            err = np.linalg.norm(res[name][k] - res["generic"][k]) / (den if den > 0 else 1.0)
            if not (err < 3e-5):
                bad += 1
                print("REVERSE MISMATCH", name, k, err, dict(n=n, batch=batch, steps=steps, nseq=nseq, knots=knots, compat=compat))
print(f"{ncase} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
