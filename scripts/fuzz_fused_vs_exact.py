"""Developer aid: randomized parity sweep of the fused path against the exact per-stage kernels (which are bit-exact to the
oracle) over grid sizes, PML widths, designs, sources, batch sizes and step counts.  `python scripts/fuzz_fused_vs_exact.py [n] [seed]`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
for k in range(cases):
    n = int(rng.choice([32, 33, 37, 48, 63, 64, 65, 70, 96, 100, 121, 128, 130, 199, 256, 257, 300, 415, 512]))
    gs = float(rng.uniform(1.0, 6.0))
    pw = float(rng.choice([0.0, 0.2, 0.5, 1.0])) * gs / 3.0
    dx = 2 * gs / (n - 1)
    dt = float(0.3 * dx / 1531.0)
    E = int(rng.choice([1, 1, 2, 3]))
    steps = int(rng.integers(3, 14))
    dim = wb.TwoDim(gs, n)
    engs = [wb.Engine(dim.x, dim.y, wb.WATER, dt, pw, 20000.0, n_env=E) for _ in range(2)]
    ts = wb.build_tspan(float(rng.uniform(0, 1e-3)), dt, steps)
    has_src, has_des = rng.random() < 0.8, rng.random() < 0.7
    u0 = (rng.standard_normal((E, 12, n, n)) * 1e-3).astype(np.float32)
    if rng.random() < 0.3:
        u0[:, 3:6] = 0
        u0[:, 9:12] = 0
    for eng in engs:
        eng.set_state(u0)
    for e in range(E):
        if has_src:
            sh = wb.build_normal(dim, [[rng.uniform(-gs, gs), rng.uniform(-gs, gs)]], [rng.uniform(0.05, 0.3) * gs], [1.0])
            for eng in engs:
                eng.set_source(sh, 1000.0, env=e)
        if has_des:
            m = int(rng.integers(1, 40))
            pos = rng.uniform(-gs, gs, (m, 2)).astype(np.float32)
            r0, r1 = rng.uniform(0.05, 0.4, m) * gs, rng.uniform(0.05, 0.4, m) * gs
            c = rng.choice([1032.0, 2120.0, 344.0], m)
            t0 = np.concatenate([pos, r0[:, None], c[:, None]], 1).astype(np.float32)
            t1 = np.concatenate([pos + rng.uniform(-0.05, 0.05, (m, 2)) * gs, r1[:, None], c[:, None]], 1).astype(np.float32)
            for eng in engs:
                eng.set_design(t0, t1, ts[0], ts[-1], env=e)
    ef, _ = engs[0].integrate(ts, wb.MODE_FUSED)
    ee, _ = engs[1].integrate(ts, wb.MODE_EXACT)
    a, b = engs[0].get_state(), engs[1].get_state()
    err = max(np.linalg.norm(a[:, f].astype(np.float64) - b[:, f]) / max(np.linalg.norm(b[:, f].astype(np.float64)), 1e-30) for f in range(12))
    eerr = np.abs(ef - ee).max() / max(ee.max(), 1e-30)
    worst = max(worst, err, eerr)
    flag = "" if (err < 2e-5 and eerr < 2e-5 and np.isfinite(a).all()) else "   <<<<<< FAIL"
    print(f"case {k:3d}: n={n:3d} gs={gs:.2f} pml={pw:.2f} E={E} steps={steps:2d} src={int(has_src)} design={int(has_des)}  field {err:.1e} energy {eerr:.1e}{flag}", flush=True)
    for eng in engs:
        eng.close()
print("worst", worst)
