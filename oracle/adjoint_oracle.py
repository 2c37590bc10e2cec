"""Oracle for the reverse pass of the Waves.jl integrator -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(waves.jl_b200/) never does.

Restates src/dynamics.jl:97-128 (`adjoint_sensitivity`, `rrule(::Integrator, ...)`) for the 2-D
AcousticDynamics (src/dynamics.jl:151-188):

  * `autograd_truth`      -- the whole unrolled trajectory in torch float64 with reverse-mode autodiff: the
                             independent check of everything below (Zygote differentiates the same program).
  * `reference_loop_torch`-- the reference's loop AS WRITTEN (src/dynamics.jl:101-115), literally: a pullback of one
                             `runge_kutta` call per stored state, applied to the running accumulator.
  * `adjoint_manual`      -- hand-derived transposed stencils (NumPy, float32 or float64), both modes; this is what
                             the CUDA kernels restate and are compared with.

Modes:
  exact  : lambda_N = a_N; lambda_i = a_i + (I + J_i^T) lambda_{i+1}        (the discrete adjoint of z_{i+1} = z_i + du(z_i))
  compat : acc = 0; for i = N..0: acc = (I + J_i^T)(acc + a_i)              (the loop as written: one extra step-vjp,
                                                                           each a_i paired with J_i instead of J_{i-1})
The parameter gradient is taken with respect to the per-cell speed plane of the total field (a perturbation that is
constant in time, accumulated over all RK stages).  The reference's own design parameters enter through a hard mask
(src/designs.jl:99-104: `<`), whose derivative is zero almost everywhere, so no reference gradient exists for them.

PARITY UNPINNED: the reference has no test or fixture for the 2-D adjoint (scripts/adjoint_sensitivity.jl is 1-D).
"""
from __future__ import annotations

import numpy as np

from . import waves_oracle as wo


# ---------------------------------------------------------------------------------------------
# torch float64 restatement of the forward step (differentiable)
# ---------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def grad_matrix_torch(g: wo.Gradient):
    torch = _torch()
    return torch.tensor(g.to_scipy().toarray().astype(np.float64))


def rhs_torch(z, c, f, D, pml, bc, c0):
    """dyn(x, t, θ) (src/dynamics.jl:179-188) for z (12, ny, nx); c (ny, nx) plane of the total field; f plane."""
    torch = _torch()
    sx, sy = pml[None, :], pml[:, None]

    def one(x, b):
        U, Vx, Vy, Px, Py, Om = x
        Vxx = Vx @ D.T            # ∂x: along the contiguous axis
        Vyy = D @ Vy              # ∂y
        Uf = U + f
        Ux, Uy = Uf @ D.T, D @ Uf
        dU = b * (Vxx + Vyy) + Px + Py - (sx + sy) * U - Om
        return torch.stack([bc * dU, Ux - sx * Vx, Uy - sy * Vy, b * sx * Vyy, b * sy * Vxx, sx * sy * U])

    return torch.cat([one(z[0:6], c * c), one(z[6:12], c0 * c0 * torch.ones_like(c))])


def rk4_du_torch(z, t, dt, cfun, ffun, D, pml, bc, c0):
    """runge_kutta (src/dynamics.jl:9-16): returns du."""
    h = 0.5 * dt
    k1 = rhs_torch(z, cfun(t), ffun(t), D, pml, bc, c0)
    k2 = rhs_torch(z + h * k1, cfun(t + h), ffun(t + h), D, pml, bc, c0)
    k3 = rhs_torch(z + h * k2, cfun(t + h), ffun(t + h), D, pml, bc, c0)
    k4 = rhs_torch(z + dt * k3, cfun(t + dt), ffun(t + dt), D, pml, bc, c0)
    return (1.0 / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4) * dt


def energy_loss_torch(traj, w_energy, dO):
    """L = sum_i sum_k w[i,k] E_k(z_i), E = (tot, inc, sc) of src/env.jl:104-111."""
    torch = _torch()
    Ut, Ui = traj[:, 0], traj[:, 6]
    E = torch.stack([(Ut ** 2).sum((1, 2)), (Ui ** 2).sum((1, 2)), ((Ut - Ui) ** 2).sum((1, 2))], 1) * dO
    return (E * w_energy).sum()


class Problem:
    """Everything the adjoint needs, in float64 NumPy: dynamics constants, per-stage speed planes, source."""

    def __init__(self, dyn: wo.AcousticDynamics, tspan, dt, speed_at, shape, freq, dO):
        self.dyn, self.tspan, self.dt = dyn, np.asarray(tspan, np.float64), float(dt)
        self.speed_at = speed_at            # t -> (ny, nx) float64 speed plane of the total field
        self.shape = None if shape is None else np.asarray(shape, np.float64)
        self.freq, self.dO = float(freq), float(dO)
        self.D = dyn.grad.to_scipy().toarray().astype(np.float64)
        self.pml = np.asarray(dyn.pml, np.float64)
        self.bc = np.asarray(dyn.bc, np.float64)
        self.c0 = float(dyn.c0)

    def source(self, t):
        if self.shape is None:
            return np.zeros_like(self.bc)
        return self.shape * np.sin(2.0 * np.pi * t * self.freq)


def autograd_truth(p: Problem, z0, w_energy, aN=None, dc=None):
    """Forward in torch float64 + autograd: (loss, dL/dz0, dL/dc, trajectory).  `dc` perturbs the speed plane at every
    stage time (used to differentiate with respect to it)."""
    torch = _torch()
    D, pml, bc = torch.tensor(p.D), torch.tensor(p.pml), torch.tensor(p.bc)
    z = torch.tensor(np.asarray(z0, np.float64), requires_grad=True)
    dcv = torch.zeros(p.bc.shape, dtype=torch.float64, requires_grad=True) if dc is None else torch.tensor(dc, requires_grad=True)
    cfun = lambda t: torch.tensor(p.speed_at(t)) + dcv
    ffun = lambda t: torch.tensor(p.source(t))
    traj = [z]
    for i in range(len(p.tspan) - 1):
        traj.append(traj[-1] + rk4_du_torch(traj[-1], p.tspan[i], p.dt, cfun, ffun, D, pml, bc, p.c0))
    T = torch.stack(traj)
    L = energy_loss_torch(T, torch.tensor(np.asarray(w_energy, np.float64)), p.dO)
    if aN is not None:
        L = L + (T[-1] * torch.tensor(np.asarray(aN, np.float64))).sum()
    gz, gc = torch.autograd.grad(L, [z, dcv])
    return float(L.detach()), gz.numpy(), gc.numpy(), T.detach().numpy()


def energy_cotangent(z, w3, dO):
    """a_i = dL/dz_i of the energy-weighted loss: only the two U planes are non-zero."""
    a = np.zeros_like(z)
    d = z[0] - z[6]
    a[0] = 2.0 * dO * (w3[0] * z[0] + w3[2] * d)
    a[6] = 2.0 * dO * (w3[1] * z[6] - w3[2] * d)
    return a


def reference_loop_torch(p: Problem, traj, w_energy, aN=None):
    """adjoint_sensitivity exactly as written (src/dynamics.jl:97-118), with torch pullbacks of one runge_kutta call."""
    torch = _torch()
    D, pml, bc = torch.tensor(p.D), torch.tensor(p.pml), torch.tensor(p.bc)
    ffun = lambda t: torch.tensor(p.source(t))
    N = len(p.tspan) - 1
    acc = np.zeros_like(traj[0])
    gc = np.zeros(p.bc.shape)
    for i in reversed(range(N + 1)):
        zi = torch.tensor(traj[i], requires_grad=True)
        dcv = torch.zeros(p.bc.shape, dtype=torch.float64, requires_grad=True)
        cfun = lambda t: torch.tensor(p.speed_at(t)) + dcv
        du = rk4_du_torch(zi, p.tspan[i], p.dt, cfun, ffun, D, pml, bc, p.c0)
        a = energy_cotangent(traj[i], w_energy[i], p.dO)
        if aN is not None and i == N:
            a = a + aN
        acc = acc + a
        gz, gci = torch.autograd.grad(du, [zi, dcv], grad_outputs=torch.tensor(acc))
        acc = acc + gz.numpy()
        gc += gci.numpy()
    return acc, gc


# ---------------------------------------------------------------------------------------------
# hand-derived reverse sweep (what the CUDA kernels restate)
# ---------------------------------------------------------------------------------------------
def _DT(g: wo.Gradient, v, axis, T):
    """(∇^T v) along `axis`: transpose of the 3-band matrix of src/operators.jl:10-22 (one-sided first / last rows)."""
    v = np.moveaxis(v, axis, -1)
    n = v.shape[-1]
    out = np.zeros_like(v)
    cm, cp = T(g.central[0]), T(g.central[1])
    # interior rows r = 1..n-2 scatter cm*v[r] to column r-1 and cp*v[r] to column r+1
    out[..., 0:n - 2] += cm * v[..., 1:n - 1]
    out[..., 2:n] += cp * v[..., 1:n - 1]
    for k in range(3):
        out[..., k] += T(g.first[k]) * v[..., 0]
        out[..., n - 3 + k] += T(g.last[k]) * v[..., n - 1]
    return np.moveaxis(out, -1, axis)


def _D(g: wo.Gradient, u, axis, T):
    u = np.moveaxis(u, axis, -1)
    out = np.empty_like(u)
    out[..., 1:-1] = T(g.central[0]) * u[..., :-2] + T(g.central[1]) * u[..., 2:]
    out[..., 0] = T(g.first[0]) * u[..., 0] + T(g.first[1]) * u[..., 1] + T(g.first[2]) * u[..., 2]
    out[..., -1] = T(g.last[0]) * u[..., -3] + T(g.last[1]) * u[..., -2] + T(g.last[2]) * u[..., -1]
    return np.moveaxis(out, -1, axis)


def rhs_T(lam, y, b_tot, dyn: wo.AcousticDynamics, T):
    """J^T lam for both wavefields and the per-cell gradient of <lam, f(y)> with respect to b = c^2 of the total field.

    (J^T λ)_U  = -(σx+σy) m λU + Dx^T λVx + Dy^T λVy + σxσy λΩ          m = bc (Dirichlet mask)
    (J^T λ)_Vx = Dx^T[b (m λU + σy λΨy)] - σx λVx
    (J^T λ)_Vy = Dy^T[b (m λU + σx λΨx)] - σy λVy
    (J^T λ)_Ψx = (J^T λ)_Ψy = m λU ;  (J^T λ)_Ω = -m λU
    g_b        = m λU (DxVx + DyVy) + σx λΨx DyVy + σy λΨy DxVx
    """
    g = dyn.grad
    sx, sy = dyn.pml.astype(T)[None, :], dyn.pml.astype(T)[:, None]
    m = dyn.bc.astype(T)
    out = np.empty_like(lam)
    gb = None
    for w, b in ((0, b_tot), (6, T(dyn.c0) * T(dyn.c0))):
        lU, lVx, lVy, lPx, lPy, lOm = (lam[w + k] for k in range(6))
        mU = m * lU
        out[w + 0] = -(sx + sy) * mU + _DT(g, lVx, -1, T) + _DT(g, lVy, -2, T) + (sx * sy) * lOm
        out[w + 1] = _DT(g, b * (mU + sy * lPy), -1, T) - sx * lVx
        out[w + 2] = _DT(g, b * (mU + sx * lPx), -2, T) - sy * lVy
        out[w + 3] = mU
        out[w + 4] = mU
        out[w + 5] = -mU
        if w == 0:
            Vxx, Vyy = _D(g, y[1], -1, T), _D(g, y[2], -2, T)
            gb = mU * (Vxx + Vyy) + (sx * lPx) * Vyy + (sy * lPy) * Vxx
    return out, gb


def rhs_fwd(y, b_tot, f, dyn: wo.AcousticDynamics, T):
    """Forward RHS in plain precision-T arithmetic (no evaluation-order claims; the bit-exact one is wo.acoustic_dynamics)."""
    g = dyn.grad
    sx, sy = dyn.pml.astype(T)[None, :], dyn.pml.astype(T)[:, None]
    m = dyn.bc.astype(T)
    out = np.empty_like(y)
    for w, b in ((0, b_tot), (6, T(dyn.c0) * T(dyn.c0))):
        U, Vx, Vy, Px, Py, Om = (y[w + k] for k in range(6))
        Vxx, Vyy = _D(g, Vx, -1, T), _D(g, Vy, -2, T)
        Uf = U + f
        out[w + 0] = m * (b * (Vxx + Vyy) + Px + Py - (sx + sy) * U - Om)
        out[w + 1] = _D(g, Uf, -1, T) - sx * Vx
        out[w + 2] = _D(g, Uf, -2, T) - sy * Vy
        out[w + 3] = (b * sx) * Vyy
        out[w + 4] = (b * sy) * Vxx
        out[w + 5] = (sx * sy) * U
    return out


def step_vjp(p: Problem, z, t, w, T):
    """(w + J_step^T w, dL/dc contribution) for one runge_kutta call at (z, t): reverse of src/dynamics.jl:9-16."""
    dt, h = T(p.dt), T(0.5 * p.dt)
    c = [p.speed_at(t).astype(T), p.speed_at(t + 0.5 * p.dt).astype(T), p.speed_at(t + p.dt).astype(T)]
    b = [ci * ci for ci in c]
    f = [p.source(t).astype(T), p.source(t + 0.5 * p.dt).astype(T)]
    k1 = rhs_fwd(z, b[0], f[0], p.dyn, T)
    y1 = z + h * k1
    k2 = rhs_fwd(y1, b[1], f[1], p.dyn, T)
    y2 = z + h * k2
    k3 = rhs_fwd(y2, b[1], f[1], p.dyn, T)
    y3 = z + dt * k3
    s6, s3 = T(p.dt / 6.0), T(p.dt / 3.0)
    ly3, g4 = rhs_T(s6 * w, y3, b[2], p.dyn, T)
    ly2, g3 = rhs_T(s3 * w + dt * ly3, y2, b[1], p.dyn, T)
    ly1, g2 = rhs_T(s3 * w + h * ly2, y1, b[1], p.dyn, T)
    lz, g1 = rhs_T(s6 * w + h * ly1, z, b[0], p.dyn, T)
    gc = T(2) * (c[0] * g1 + c[1] * (g2 + g3) + c[2] * g4)
    return w + ((ly3 + ly2) + ly1) + lz, gc


def adjoint_manual(p: Problem, traj, w_energy, aN=None, mode="exact", dtype=np.float64):
    """Reverse sweep over a stored trajectory (steps+1, 12, ny, nx): returns (dL/dz0, dL/dc)."""
    T = np.dtype(dtype).type
    N = len(p.tspan) - 1
    traj = np.asarray(traj, dtype)
    cot = lambda i: energy_cotangent(traj[i], np.asarray(w_energy[i], dtype), T(p.dO)) + (0 if (aN is None or i != N) else np.asarray(aN, dtype))
    gc = np.zeros(p.bc.shape, dtype)
    if mode == "exact":
        lam = cot(N)
        for i in reversed(range(N)):
            v, g = step_vjp(p, traj[i], p.tspan[i], lam, T)
            lam = cot(i) + v
            gc += g
        return lam, gc
    if mode == "compat":
        acc = np.zeros_like(traj[0])
        for i in reversed(range(N + 1)):
            acc, g = step_vjp(p, traj[i], p.tspan[i], acc + cot(i), T)
            gc += g
        return acc, gc
    raise ValueError(mode)
