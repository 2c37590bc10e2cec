"""Developer diagnostic (run on a GPU box): CUDA paths vs the C oracle with per-field error maps."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waves_b200 as wb  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import waves_oracle as wo  # noqa: E402

NAMES = ["U", "Vx", "Vy", "Px", "Py", "Om"]


def report(tag, got, ref):
    bad = False
    for f in range(12):
        d = got[f].astype(np.float64) - ref[f]
        nrm = np.linalg.norm(ref[f]) + 1e-300
        rel = np.linalg.norm(d) / nrm
        if not np.isfinite(rel) or rel > 1e-5:
            j, i = np.unravel_index(np.argmax(np.abs(d)), d.shape)
            nb = int((np.abs(d) > 1e-4 * (np.abs(ref[f]).max() + 1e-30)).sum())
            rows = np.unique(np.nonzero(np.abs(d) > 1e-4 * (np.abs(ref[f]).max() + 1e-30))[0])
            cols = np.unique(np.nonzero(np.abs(d) > 1e-4 * (np.abs(ref[f]).max() + 1e-30))[1])
            print(f"  {tag} {'tot' if f < 6 else 'inc'}.{NAMES[f % 6]}: rel {rel:.3e} max|d| {np.abs(d).max():.3e} at (j={j}, i={i}) "
                  f"bad cells {nb} rows[{rows[:6]}..{rows[-3:] if len(rows) else ''}] cols[{cols[:6]}..{cols[-3:] if len(cols) else ''}]")
            bad = True
    tot = np.linalg.norm(got.astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-300)
    print(f"{tag}: overall rel-L2 {tot:.3e} bitwise={np.array_equal(got, ref)} {'<-- BAD' if bad else 'ok'}")
    return tot


def case(n, gs, pmlw, steps, design, seed=0, random_state=True, src_mu=(-1.0, 0.2), src_sigma=0.15, t0=3e-4, dt=1e-5):
    dim = wo.TwoDim.make(gs, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, pmlw, 20000.0)
    grid = wo.build_grid(dim)
    shape = wo.build_normal(grid, np.array([src_mu]), np.array([src_sigma]), np.array([1.0]))
    rng = np.random.default_rng(seed)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(np.float32) if random_state else np.zeros((12, n, n), np.float32)
    tspan = wo.build_tspan(np.float32(t0), np.float32(dt), steps)
    dO = np.float32(wo.get_dx(dim) * wo.get_dy(dim))
    d0, d1 = design if design else (None, None)
    t = time.time()
    ref, ren, rfr = co.integrate(dyn, u0, tspan, dt, dO, d0, d1, tspan[0], tspan[-1], shape=shape, freq=1000.0, save_steps=[1])
    t_or = time.time() - t
    out = {}
    for mode, name in ((wb.MODE_EXACT, "exact"), (wb.MODE_FUSED, "fused")):
        eng = wb.Engine(dim.x, dim.y, wo.WATER, dt, pmlw, 20000.0, n_env=1, sigma=dyn.pml, grad8=co.grad8(dyn.grad), d_omega=float(dO))
        eng.set_state(u0[None])
        eng.set_source(shape, 1000.0)
        if d0 is not None:
            eng.set_design(co._design_args(d0, d1)[1], co._design_args(d0, d1)[2], tspan[0], tspan[-1])
        if mode == wb.MODE_EXACT:
            k = eng.rhs(tspan[1])
            kr = co.rhs(dyn, u0, tspan[1], d0, d1, tspan[0], tspan[-1], shape=shape, freq=1000.0)
            report(f"[n={n}] rhs exact", k, kr)
        t = time.time()
        en, fr = eng.integrate(tspan, mode, energy=True, save_steps=[1])
        dtm = time.time() - t
        got = eng.get_state(0)
        report(f"[n={n}] {name} step1", fr[0, 0], rfr[0])
        r = report(f"[n={n}] {name} final({steps})", got, ref)
        erel = np.abs(en[0] - ren).max() / (np.abs(ren).max() + 1e-30)
        print(f"[n={n}] {name}: energy rel {erel:.3e} E_end gpu {en[0, -1]} oracle {ren[-1]}  time {dtm*1e3:.1f} ms (oracle {t_or*1e3:.0f} ms)")
        out[name] = r
        eng.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "ragged", "c1"]
    pos = np.array([[0.5, 0.0], [0.9, 0.3], [-0.2, -1.1]], dtype=np.float32)
    des = (wo.Cylinders(pos, [0.4, 0.3, 0.5], [1032.0, 1032.0, 2120.0]), wo.Cylinders(pos, [0.6, 0.25, 0.35], [1032.0, 1032.0, 2120.0]))
    if "small" in which:
        case(96, 3.0, 0.6, 40, des)
        case(96, 3.0, 0.6, 8, None)
    if "ragged" in which:
        case(70, 2.0, 0.5, 12, des, seed=3)
        case(33, 1.0, 0.3, 12, None, seed=4, dt=5e-6)
        case(257, 6.0, 1.0, 12, des, seed=5)
    if "c1" in which:
        case(700, 15.0, 2.0, 20, None, random_state=False, src_mu=(-10.0, 0.0), src_sigma=0.3, t0=0.0)
        ds = wo.build_triple_ring_design_space()
        rng = np.random.default_rng(0)
        d0 = ds.sample(rng)
        d1 = ds(d0, wo.build_action_space(d0, 0.25).sample(rng))
        case(700, 15.0, 2.0, 10, (d0, d1), random_state=True, src_mu=(-3.0, 1.0), src_sigma=0.3, t0=1e-3)
