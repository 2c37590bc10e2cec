"""CPU tests of the adjoint oracle: the hand-derived reverse sweep (what the CUDA kernels restate) against torch
autograd over the unrolled trajectory, and the reference-loop-as-written (src/dynamics.jl:97-118) against its
literal torch transcription."""
import numpy as np

from oracle import adjoint_oracle as ao
from oracle import waves_oracle as wo


def make_problem(n=24, steps=5, moving=True):
    dim = wo.TwoDim.make(1.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.3, 20000.0)
    grid = wo.build_grid(dim)
    shape = wo.build_normal(grid, np.array([[-0.4, 0.1]]), np.array([0.12]), np.array([1.0]))
    pos = np.array([[0.2, 0.0], [-0.1, -0.4]], np.float32)
    d0 = wo.Cylinders(pos, [0.25, 0.2], [1032.0, 2100.0])
    d1 = wo.Cylinders(pos, [0.35, 0.15] if moving else [0.25, 0.2], [1032.0, 2100.0])
    dt = 2e-6
    tspan = np.float64(wo.build_tspan(np.float32(1e-4), np.float32(dt), steps))
    interp = wo.DesignInterpolator(d0, d1, np.float32(tspan[0]), np.float32(tspan[-1]))
    speed_at = lambda t: np.asarray(wo.speed(interp(np.float32(t)), grid, dyn.c0), np.float64)
    dO = float(wo.get_dx(dim) * wo.get_dy(dim))
    p = ao.Problem(dyn, tspan, dt, speed_at, shape, 1000.0, dO)
    rng = np.random.default_rng(3)
    z0 = rng.standard_normal((12, n, n)) * 1e-3
    w = rng.uniform(0.2, 1.0, (steps + 1, 3))
    aN = rng.standard_normal((12, n, n)) * 1e-4
    return p, z0, w, aN


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_manual_exact_adjoint_matches_autograd():
    p, z0, w, aN = make_problem()
    L, gz, gc, traj = ao.autograd_truth(p, z0, w, aN)
    mz, mc = ao.adjoint_manual(p, traj, w, aN, "exact", np.float64)
    assert rel(mz, gz) < 1e-11 and rel(mc, gc) < 1e-11
    # float32 sweep stays within the float32 budget of the north star (1e-4)
    fz, fc = ao.adjoint_manual(p, traj, w, aN, "exact", np.float32)
    assert rel(fz, gz) < 1e-4 and rel(fc, gc) < 1e-4


def test_manual_compat_matches_reference_loop_as_written():
    p, z0, w, aN = make_problem(steps=4)
    _, gz, gc, traj = ao.autograd_truth(p, z0, w, aN)
    rz, rc = ao.reference_loop_torch(p, traj, w, aN)
    mz, mc = ao.adjoint_manual(p, traj, w, aN, "compat", np.float64)
    assert rel(mz, rz) < 1e-11 and rel(mc, rc) < 1e-11
    # the loop as written is NOT the exact discrete adjoint: it applies one extra step-vjp
    assert rel(rz, gz) > 1e-3


def test_autograd_matches_finite_differences():
    p, z0, w, aN = make_problem(n=20, steps=3)
    L, gz, gc, _ = ao.autograd_truth(p, z0, w, aN)
    rng = np.random.default_rng(0)
    for _ in range(3):
        dz = rng.standard_normal(z0.shape) * 1e-3
        eps = 1e-4
        Lp = ao.autograd_truth(p, z0 + eps * dz, w, aN)[0]
        Lm = ao.autograd_truth(p, z0 - eps * dz, w, aN)[0]
        fd = (Lp - Lm) / (2 * eps)
        assert abs(fd - np.sum(gz * dz)) <= 1e-6 * abs(fd) + 1e-18
    dc = rng.standard_normal(p.bc.shape)
    eps = 1e-3
    Lp = ao.autograd_truth(p, z0, w, aN, dc=eps * dc)[0]
    Lm = ao.autograd_truth(p, z0, w, aN, dc=-eps * dc)[0]
    fd = (Lp - Lm) / (2 * eps)
    assert abs(fd - np.sum(gc * dc)) <= 1e-5 * abs(fd) + 1e-18


def test_transposed_operator_dot_product():
    """<J v, w> == <v, J^T w> for the hand-derived transpose, including one-sided rows, PML and the mask."""
    p, z0, _, _ = make_problem()
    rng = np.random.default_rng(1)
    T = np.float64
    b = p.speed_at(p.tspan[0]) ** 2
    v, w = rng.standard_normal(z0.shape), rng.standard_normal(z0.shape)
    f0 = np.zeros_like(p.bc)
    Jv = ao.rhs_fwd(v, b, f0, p.dyn, T)         # the RHS is linear in the state when f = 0
    JTw, _ = ao.rhs_T(w, v, b, p.dyn, T)
    assert abs(np.sum(Jv * w) - np.sum(v * JTw)) < 1e-10 * abs(np.sum(Jv * w))
