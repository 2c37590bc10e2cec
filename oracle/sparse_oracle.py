"""Second, structurally independent CPU statement of the reference's 2-D hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product (waves.jl_b200/) never does.

oracle/waves_oracle.py restates the path with stencil slices and fused expressions; this file states it the way the reference's
source reads, line for line, in the reference's own array layout, so that the two statements share no code and no design:

  * arrays are (nx, ny[, field]) like Julia's (dim 1 = x), every operation is its own float32 NumPy broadcast (one rounding
    per reference operation, no fusion);
  * the derivative IS a sparse matrix: `gradient(x)` builds the dense coefficient matrix exactly as src/operators.jl:10-22 does
    and converts it to SciPy CSC (SparseArrays is CSC), and  ∂x(∇, u) = ∇ * u,  ∂y(∇, u) = (∇ * u')'  (src/operators.jl:45-46)
    are SciPy sparse-times-dense products (column-ordered accumulation, like SparseArrays' spmm);
  * `acoustic_dynamics`, the TwoDim call of AcousticDynamics, `runge_kutta`, the Integrator loop, `build_pml`, `build_normal`,
    `location_mask`, `speed`, DesignInterpolator and the energy metric follow src/dynamics.jl:9-16,37-49,151-188,
    src/pml.jl:21-29, src/utils.jl:12-18, src/designs.jl:99-116,287-292 and src/env.jl:104-111 statement by statement.

tests/test_oracle.py::test_sparse_statement_* require float32 agreement of the two statements (<= 1e-6 relative, i.e. the
order-of-accumulation noise of a handful of float32 operations) on BASELINE config 1 (700^2, 100 steps: fields at probes and
the 101 x 3 energy trace, against the committed golden fixture) and on the 96^2 moving-design fixture (all 12 fields).
PARITY UNPINNED with respect to the real reference (Julia is not installed; see oracle/waves_oracle.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

F32 = np.float32
FORWARD_DIFF_COEF = np.array([-3.0, 4.0, -1.0], F32)    # src/operators.jl:3
BACKWARD_DIFF_COEF = np.array([1.0, -4.0, 3.0], F32)    # src/operators.jl:4
CENTRAL_DIFF_COEF = np.array([-1.0, 1.0], F32)          # src/operators.jl:5


def jl_range(start, stop, n):
    """collect(range(start, stop, n)) for Float32 endpoints (src/dims.jl:56-60): Julia evaluates the range in twice precision
    and rounds every element once; float64 interpolation of the float32 endpoints reproduces that for these sizes."""
    a, b = np.float64(F32(start)), np.float64(F32(stop))
    i = np.arange(n, dtype=np.float64)
    return (((n - 1 - i) * a + i * b) / (n - 1)).astype(F32)


def gradient(x):
    """src/operators.jl:10-22."""
    n = x.shape[0]
    grad = np.zeros((n, n), F32)
    delta = F32(F32(x[-1] - x[0]) / F32(n - 1))
    grad[[0, 1, 2], 0] = FORWARD_DIFF_COEF
    grad[[n - 3, n - 2, n - 1], n - 1] = BACKWARD_DIFF_COEF
    for i in range(1, n - 1):
        grad[[i - 1, i + 1], i] = CENTRAL_DIFF_COEF
    return sp.csc_matrix((grad / F32(F32(2.0) * delta)).T.astype(F32))


def dx_(G, u):   # ∂x(∇, u) = ∇ * u            src/operators.jl:45
    return np.asarray(G @ u, F32)


def dy_(G, u):   # ∂y(∇, u) = (∇ * u')'        src/operators.jl:46
    return np.asarray(G @ np.ascontiguousarray(u.T), F32).T


def build_pml(x, ny, width, scale):
    """src/pml.jl:21-29 -> (nx, ny)."""
    x = np.abs(x).astype(F32)
    pml_start = F32(x[0] - F32(width))
    region = x > pml_start
    x[~region] = F32(0.0)
    x[region] = (x[region] - x[region].min()) / F32(width)
    x = np.repeat(x[:, None], ny, axis=1)
    return (((x * x) * x) * F32(scale)).astype(F32)   # `x .^ 3`: Base.literal_pow(^, x, Val(3)) == x*x*x for Float32


def build_dirichlet(nx, ny):
    """src/dims.jl:117-124."""
    bc = np.ones((nx, ny), F32)
    bc[:, 0] = 0
    bc[0, :] = 0
    bc[:, -1] = 0
    bc[-1, :] = 0
    return bc


def build_grid(x, y):
    """src/dims.jl:92-97 -> (nx, ny, 2)."""
    gx = np.repeat(x[:, None], len(y), axis=1)
    gy = np.repeat(y[None, :], len(x), axis=0)
    return np.stack([gx, gy], axis=2).astype(F32)


def build_normal(grid, mu, sigma, a):
    """src/utils.jl:12-18: sum_k a_k / (2π σ_k²) exp(-|x - μ_k|² / (2 σ_k²)); exp evaluated like Julia's exp(::Float32)
    (correctly rounded from a double evaluation)."""
    out = np.zeros(grid.shape[:2], F32)
    for k in range(len(sigma)):
        s, ak = F32(sigma[k]), F32(a[k])
        d = grid - np.asarray(mu[k], F32)[None, None, :]
        r2 = (d[:, :, 0] ** 2 + d[:, :, 1] ** 2).astype(F32)
        coef = F32(F32(1.0) / F32(F32(F32(2.0) * F32(np.pi)) * F32(s ** 2)))
        e = np.exp((-r2 / F32(F32(2.0) * F32(s ** 2))).astype(np.float64)).astype(F32)
        out = out + F32(coef * ak) * e if k else (F32(coef * ak) * e).astype(F32)
    return out.astype(F32)


def speed(pos, r, c, grid, ambient):
    """location_mask + speed, src/designs.jl:99-116."""
    d = grid[:, :, None, :] - np.asarray(pos, F32)[None, None, :, :]
    mask = (d[..., 0] ** 2 + d[..., 1] ** 2).astype(F32) < (np.asarray(r, F32) ** 2)[None, None, :]
    ambient_mask = mask.sum(axis=2) == 0
    C0 = ambient_mask.astype(F32) * F32(ambient)
    C_design = np.zeros(grid.shape[:2], F32)
    for k in range(mask.shape[2]):   # sum(...; dims = 3) over the cylinders, in order
        C_design = C_design + mask[:, :, k].astype(F32) * F32(c[k])
    return (C0 + C_design).astype(F32)


def design_at(t, d0, d1, ti, tf):
    """(interp::DesignInterpolator)(t), src/designs.jl:287-292, for (ncyl, 4) tables {x, y, r, c}:
    initial + (clamp(t, ti, tf) - ti) * ((final + initial * -1) / Δt)   (Cylinders `-` is `+` of `* -1.0f0`, `/` is `* (1/n)`)."""
    ti, tf, t = F32(ti), F32(tf), F32(t)
    dt = F32(tf - ti)
    dt = dt if dt > 0 else F32(1.0)
    dy = (d1 + d0 * F32(-1.0)).astype(F32)
    slope = (dy * F32(F32(1.0) / dt)).astype(F32)
    return (d0 + slope * F32(min(max(t, ti), tf) - ti)).astype(F32)


def acoustic_dynamics(x, c, f, G, pml, bc):
    """src/dynamics.jl:151-177, statement by statement; x (nx, ny, 6)."""
    U, Vx, Vy, Px, Py, Om = (x[:, :, k] for k in range(6))
    b = (c * c).astype(F32) if isinstance(c, np.ndarray) else F32(F32(c) * F32(c))
    sx = pml
    sy = sx.T
    Vxx = dx_(G, Vx)
    Vyy = dy_(G, Vy)
    Uf = (U + f).astype(F32)
    Ux = dx_(G, Uf)
    Uy = dy_(G, Uf)
    dU = (b * (Vxx + Vyy)).astype(F32)
    dU = dU + Px
    dU = dU + Py
    dU = dU - (sx + sy) * U
    dU = dU - Om
    dVx = Ux - sx * Vx
    dVy = Uy - sy * Vy
    dPx = (b * sx) * Vyy
    dPy = (b * sy) * Vxx
    dOm = (sx * sy) * U
    return np.stack([bc * dU, dVx, dVy, dPx, dPy, dOm], axis=2).astype(F32)


class Dynamics:
    """AcousticDynamics(dim, c0, pml_width, pml_scale), src/dynamics.jl:141-149, and its TwoDim call :179-188."""

    def __init__(self, grid_size, n, c0, pml_width, pml_scale):
        self.x = jl_range(-F32(grid_size), F32(grid_size), n)
        self.y = self.x.copy()
        self.c0 = F32(c0)
        self.grad = gradient(self.x)
        self.pml = build_pml(self.x, n, pml_width, pml_scale)
        self.bc = build_dirichlet(n, n)
        self.grid = build_grid(self.x, self.y)

    def __call__(self, x, t, theta):
        C, F = theta
        c, f = C(t), F(t)
        dtot = acoustic_dynamics(x[:, :, 0:6], c, f, self.grad, self.pml, self.bc)
        dinc = acoustic_dynamics(x[:, :, 6:12], self.c0, f, self.grad, self.pml, self.bc)
        return np.concatenate([dtot, dinc], axis=2)


def source(shape, freq):
    """(source)(t) = shape .* sin.(2.0f0 * pi * t * freq), src/sources.jl:67-69: float32 argument, Julia's sin(::Float32)."""
    def F(t):
        arg = F32(F32(F32(F32(2.0) * F32(np.pi)) * F32(t)) * F32(freq))
        return (shape * F32(np.sin(np.float64(arg)))).astype(F32)
    return F


def runge_kutta(f, u, t, theta, dt):
    """src/dynamics.jl:9-16."""
    dt = F32(dt)
    h = F32(F32(0.5) * dt)
    k1 = f(u, t, theta)
    k2 = f((u + h * k1).astype(F32), F32(t + h), theta)
    k3 = f((u + h * k2).astype(F32), F32(t + h), theta)
    k4 = f((u + dt * k3).astype(F32), F32(t + dt), theta)
    du = F32(F32(1.0) / F32(6.0)) * (((k1 + F32(2.0) * k2) + F32(2.0) * k3) + k4)
    return (du * dt).astype(F32)


def integrate(dyn, ui, tspan, theta, dt):
    """(iter::Integrator)(ui, tspan, θ), src/dynamics.jl:37-49 -> list of states, ui first."""
    sol = [np.asarray(ui, F32)]
    for i in range(len(tspan) - 1):
        sol.append((sol[-1] + runge_kutta(dyn, sol[-1], F32(tspan[i]), theta, dt)).astype(F32))
    return sol


def energies(sol, dx, dy):
    """src/env.jl:104-111 -> (frames, 3); float64 accumulation (the reference's float32 sum order is unspecified)."""
    dO = F32(F32(dx) * F32(dy))
    out = np.empty((len(sol), 3), F32)
    for i, z in enumerate(sol):
        ut, ui = z[:, :, 0].astype(np.float64), z[:, :, 6].astype(np.float64)
        out[i] = [F32(np.sum(ut ** 2)) * dO, F32(np.sum(ui ** 2)) * dO, F32(np.sum((z[:, :, 0] - z[:, :, 6]).astype(np.float64) ** 2)) * dO]
    return out


def to_julia(state_c):
    """(12, ny, nx) C-ordered state of the other oracle / the CUDA library -> (nx, ny, 12)."""
    return np.ascontiguousarray(np.transpose(state_c, (2, 1, 0)))


def from_julia(state_j):
    return np.ascontiguousarray(np.transpose(state_j, (2, 1, 0)))
