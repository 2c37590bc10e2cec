"""GPU parity of the reverse pass (waves_adjoint; src/dynamics.jl:97-128) against the adjoint oracle: torch autograd over
the unrolled float64 trajectory (exact mode) and the reference loop as written (compat mode)."""
import numpy as np
import pytest

import waves_b200 as wb
from oracle import adjoint_oracle as ao
from oracle import c_oracle as co
from oracle import waves_oracle as wo

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north-star float32 tolerance (relative L2)


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b)


def setup(n=48, steps=6, moving=True, pml_width=0.5):
    dim = wo.TwoDim.make(2.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, pml_width, 20000.0)
    grid = wo.build_grid(dim)
    shape = wo.build_normal(grid, np.array([[-0.8, 0.2]]), np.array([0.2]), np.array([1.0]))
    pos = np.array([[0.4, 0.0], [-0.2, -0.7]], np.float32)
    d0 = wo.Cylinders(pos, [0.45, 0.3], [1032.0, 2100.0])
    d1 = wo.Cylinders(pos, [0.55, 0.25] if moving else [0.45, 0.3], [1032.0, 2100.0])
    dt = 4e-6
    ts32 = wo.build_tspan(np.float32(2e-4), np.float32(dt), steps)
    interp = wo.DesignInterpolator(d0, d1, ts32[0], ts32[-1])
    speed_at = lambda t: np.asarray(wo.speed(interp(np.float32(t)), grid, dyn.c0), np.float64)
    dO = np.float32(wo.get_dx(dim) * wo.get_dy(dim))
    p = ao.Problem(dyn, np.float64(ts32), dt, speed_at, shape, 1000.0, float(dO))
    rng = np.random.default_rng(5)
    z0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(np.float32)
    w = rng.uniform(0.2, 1.0, (steps + 1, 3)).astype(np.float32)
    aN = (rng.standard_normal((12, n, n)) * 1e-4).astype(np.float32)
    eng = wb.Engine(dim.x, dim.y, dyn.c0, dt, sigma=dyn.pml, grad8=co.grad8(dyn.grad), d_omega=float(dO))
    tab = lambda c: np.concatenate([c.pos, c.r[:, None], c.c[:, None]], 1).astype(np.float32)
    eng.set_source(shape, 1000.0)
    eng.set_design(tab(d0), tab(d1), ts32[0], ts32[-1])
    return p, eng, ts32, z0, w, aN


@pytest.mark.parametrize("fwd", [wb.MODE_EXACT, wb.MODE_FUSED])
def test_exact_adjoint_matches_autograd(fwd):
    p, eng, ts, z0, w, aN = setup()
    L, gz, gc, traj = ao.autograd_truth(p, z0, w, aN)
    eng.set_state(z0[None])
    loss, dz0, dc = eng.adjoint(ts, w, aN[None], fwd_mode=fwd, adj_mode=wb.ADJ_EXACT)
    assert rel(dz0[0], gz) < TOL and rel(dc[0], gc) < TOL
    L_energy = L - float(np.sum(traj[-1] * aN))
    assert abs(loss[0] - L_energy) < 1e-4 * abs(L_energy)
    # the handle is left at z_N like the forward call
    assert rel(eng.get_state(0), traj[-1]) < TOL
    eng.close()


def test_compat_mode_is_the_reference_loop_as_written():
    p, eng, ts, z0, w, aN = setup(steps=5)
    _, gz, _, traj = ao.autograd_truth(p, z0, w, aN)
    rz, rc = ao.reference_loop_torch(p, traj, w, aN)
    eng.set_state(z0[None])
    _, dz0, dc = eng.adjoint(ts, w, aN[None], fwd_mode=wb.MODE_EXACT, adj_mode=wb.ADJ_COMPAT)
    assert rel(dz0[0], rz) < TOL and rel(dc[0], rc) < TOL
    assert rel(dz0[0], gz) > 1e-3   # and it is not the exact discrete adjoint
    eng.close()


def test_adjoint_without_design_and_final_state_cotangent_only():
    """NoDesign: scalar c0 everywhere; loss = <aN, z_N> only (the generic rrule use)."""
    p, eng, ts, z0, w, aN = setup(moving=False)
    eng.set_design(None, None, 0, 0)
    grid_c = np.full(p.bc.shape, p.c0)
    p.speed_at = lambda t: grid_c
    _, gz, gc, _ = ao.autograd_truth(p, z0, np.zeros_like(w), aN)
    eng.set_state(z0[None])
    _, dz0, dc = eng.adjoint(ts, None, aN[None], fwd_mode=wb.MODE_FUSED)
    assert rel(dz0[0], gz) < TOL and rel(dc[0], gc) < TOL
    eng.close()


def test_adjoint_is_linear_in_the_loss_weights_at_full_size():
    """700^2, 20 steps: dL/dz0 for weights (w1 + w2) equals the sum of the two sweeps (size-independent property)."""
    dim = wb.TwoDim(15.0, 700)
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0)
    eng.set_source(wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0]), 1000.0)
    rng = np.random.default_rng(0)
    ds = wb.build_triple_ring_design_space()
    d0 = ds.rand(rng)
    ts = wb.build_tspan(0.0, 1e-5, 20)
    eng.set_design(d0.table(), d0.table(), ts[0], ts[-1])
    z0 = (rng.standard_normal((1, 12, 700, 700)) * 1e-3).astype(np.float32)
    w1 = np.zeros((21, 3), np.float32); w1[-1, 2] = 1.0
    w2 = np.zeros((21, 3), np.float32); w2[10, 0] = 0.5
    out = []
    for w in (w1, w2, w1 + w2):
        eng.set_state(z0)
        out.append(eng.adjoint(ts, w))
    assert rel(out[0][1] + out[1][1], np.float64(out[2][1])) < 1e-5
    assert rel(out[0][2] + out[1][2], np.float64(out[2][2])) < 1e-5
    assert abs(out[0][0] + out[1][0] - out[2][0]) < 1e-5 * abs(out[2][0])
    eng.close()
