"""Generates the golden fixtures in tests/golden/ from the NumPy oracle.

The reference (Julia) cannot run in this image and holds no golden vectors for
the 2-D path, so these fixtures pin the *restatement* (oracle/waves_oracle.py):
the C oracle and the CUDA path are both checked against them.  PARITY UNPINNED
with respect to the real reference -- see oracle/waves_oracle.py header.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co  # noqa: E402
from oracle import waves_oracle as wo  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small_case():
    """96^2 grid, 3 cylinders (two overlapping -> speeds sum), moving radii, source, PML, 40 steps."""
    n = 96
    dim = wo.TwoDim.make(3.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.6, 20000.0)
    grid = wo.build_grid(dim)
    shape = wo.build_normal(grid, np.array([[-1.0, 0.2]]), np.array([0.15]), np.array([1.0]))
    pos = np.array([[0.5, 0.0], [0.9, 0.3], [-0.2, -1.1]], dtype=np.float32)
    d0 = wo.Cylinders(pos, [0.4, 0.3, 0.5], [1032.0, 1032.0, 2120.0])
    d1 = wo.Cylinders(pos, [0.6, 0.25, 0.35], [1032.0, 1032.0, 2120.0])
    rng = np.random.default_rng(7)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(np.float32)
    dt = wo.F32(1e-5)
    tspan = wo.build_tspan(wo.F32(3e-4), dt, 40)
    interp = wo.DesignInterpolator(d0, d1, tspan[0], tspan[-1])
    src = wo.Source(shape, wo.F32(1000.0))
    theta = (lambda t: wo.speed(interp(t), grid, dyn.c0), src)
    k0 = dyn(u0, tspan[5], theta)
    sol = wo.integrate(dyn, u0, tspan, theta, dt)
    en = wo.energies(sol[:, 0], sol[:, 6], dim)
    dO = wo.F32(wo.get_dx(dim) * wo.get_dy(dim))
    # the C restatement must agree bit for bit
    st, enc, fr = co.integrate(dyn, u0, tspan, dt, dO, d0, d1, tspan[0], tspan[-1], shape=shape, freq=1000.0,
                               save_steps=[1])
    assert np.array_equal(st, sol[-1]) and np.array_equal(fr[0], sol[1])
    np.savez_compressed(
        os.path.join(HERE, "small_design_96.npz"), n=n, grid_size=np.float32(3.0), pml_width=np.float32(0.6),
        pml_scale=np.float32(20000.0), c0=wo.WATER, dt=dt, tspan=tspan, u0=u0, shape=shape, freq=np.float32(1000.0),
        cyl0=np.concatenate([d0.pos, d0.r[:, None], d0.c[:, None]], 1),
        cyl1=np.concatenate([d1.pos, d1.r[:, None], d1.c[:, None]], 1), rhs_t=tspan[5], rhs=k0, step1=sol[1],
        final=sol[-1], energy=en, x=dim.x, sigma=dyn.pml, grad8=co.grad8(dyn.grad), dOmega=dO)
    print("small_design_96: E_final", en[-1])


def config1():
    """BASELINE config 1: TwoDim(15,700), Gaussian source at (-10,0), no design, 100 steps."""
    n = 700
    dim = wo.TwoDim.make(15.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 2.0, 20000.0)
    grid = wo.build_grid(dim)
    shape = wo.build_normal(grid, np.array([[-10.0, 0.0]]), np.array([0.3]), np.array([1.0]))
    dt = wo.F32(1e-5)
    tspan = wo.build_tspan(wo.F32(0.0), dt, 100)
    src = wo.Source(shape, wo.F32(1000.0))
    theta = (lambda t: dyn.c0, src)
    u = np.zeros((12, n, n), dtype=np.float32)
    rng = np.random.default_rng(1)
    # probes: all 12 fields at 400 points around the source + 112 anywhere
    pj = np.concatenate([rng.integers(250, 450, 400), rng.integers(0, n, 112)])
    pi = np.concatenate([rng.integers(20, 220, 400), rng.integers(0, n, 112)])
    keep = {1: None, 10: None, 100: None}
    U_tot, U_inc = [u[0].copy()], [u[6].copy()]
    sums = {}
    for i in range(100):
        u = u + wo.runge_kutta(dyn, u, tspan[i], theta, dt)
        U_tot.append(u[0].copy())
        U_inc.append(u[6].copy())
        if (i + 1) in keep:
            keep[i + 1] = u[:, pj, pi].copy()
            sums[i + 1] = np.array([np.sum(u[f].astype(np.float64)) for f in range(12)])
    en = wo.energies(np.stack(U_tot), np.stack(U_inc), dim)
    dO = wo.F32(wo.get_dx(dim) * wo.get_dy(dim))
    st, enc, _ = co.integrate(dyn, np.zeros((12, n, n), np.float32), tspan, dt, dO, shape=shape, freq=1000.0)
    assert np.array_equal(st, u), "C oracle != NumPy oracle"
    np.savez_compressed(
        os.path.join(HERE, "config1_700.npz"), n=n, tspan=tspan, energy=en, probe_j=pj, probe_i=pi,
        probes_1=keep[1], probes_10=keep[10], probes_100=keep[100], sums_1=sums[1], sums_10=sums[10],
        sums_100=sums[100], x=dim.x, sigma=dyn.pml, grad8=co.grad8(dyn.grad), dOmega=dO,
        shape_row350=shape[350].copy(), shape_sum=np.float64(shape.astype(np.float64).sum()))
    print("config1_700: E[100]", en[100], "E[50]", en[50])


if __name__ == "__main__":
    small_case()
    config1()
