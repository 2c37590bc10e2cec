"""CPU oracle for the Waves.jl 1-D latent dynamics -- TEST INFRASTRUCTURE ONLY.

NumPy restatement, op for op and in IEEE float32, of the batched one-dimensional acoustic
dynamics the reference's AcousticEnergyModel integrates at training time:

    (dyn::AcousticDynamics{OneDim})(x, t, θ)        src/dynamics.jl:190-222
    runge_kutta / (iter::Integrator)(ui, tspan, θ)  src/dynamics.jl:9-16, :37-49
    LinearInterpolation / linear_interp             src/utils.jl:70-98     (θ[1] = C)
    Source(shape, freq)(t::Vector)                  src/sources.jl:10-23   (θ[2] = F)
    build_pml(::OneDim), build_dirichlet(::OneDim)  src/pml.jl:6-15, src/dims.jl:111-115
    compute_latent_energy                           src/model/acoustic_energy_model.jl:6-15
    adjoint_sensitivity (batchwise OneDim)          src/dynamics.jl:97-118 (see latent_adjoint_* below)

It is the *checker* of `waves_latent_integrate` / `waves_latent_adjoint`: only `tests/` may import it.

PARITY UNPINNED: the reference cannot run here (Julia) and holds no golden vectors for this path.
`test/pinn.jl:20-44` restates the same 1-D right-hand side independently (SimpleWave) and is what
the form of the equations is checked against.

Array convention: reference arrays are column-major; the SAME memory image in C order is used here:
    state   (n, 4, batch)          -> [batch][4][n]          fields U_tot, V_tot, U_inc, V_inc
    z       (n, 4, batch, time)    -> [time][batch][4][n]
    tspan   (time, batch)          -> [batch][time]
    C.X     (nseq, batch)          -> [batch][nseq]
    C.Y     (n, nseq, batch)       -> [batch][nseq][n]
    F.shape (n, batch), PML (n, batch) -> [batch][n]
    energy  (time, 3, batch)       -> [batch][3][time]

One version-dependent spot: `dyn.c0 * ∇ * (U_inc .+ f)` (src/dynamics.jl:213) is the 3-argument `*`.
On Julia >= 1.7 LinearAlgebra evaluates `*(α::Number, A::AbstractMatrix, B)` for a sparse A as
`(A*B) .* α`; older versions as `(α*A)*B`.  The reference pins no Julia version; the first form is
restated (the two differ by an ulp of the product).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .waves_oracle import F32, Gradient, OneDim, apply_gradient, build_gradient, get_dx, source_factor  # noqa: F401


def build_pml_1d(x: np.ndarray, width, scale) -> np.ndarray:
    """build_pml(::OneDim, width, scale), src/pml.jl:6-15."""
    T = x.dtype.type
    ax = np.abs(x)
    start = T(min(ax[0], ax[-1]) - T(width))
    pml = np.maximum(ax - start, T(0)) / T(width)
    pml = np.clip(pml, T(0), T(1))
    return ((pml * pml) * pml) * T(scale)          # pml .^ 3 * scale (literal power: x*x*x)


def build_dirichlet_1d(n: int, dtype=F32) -> np.ndarray:
    """build_dirichlet(::OneDim), src/dims.jl:111-115."""
    bc = np.ones(n, dtype=dtype)
    bc[[0, -1]] = 0
    return bc


def linear_interp(X: np.ndarray, Y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """linear_interp(X, Y, x), src/utils.jl:70-86.  X [batch][nseq], Y [batch][nseq][n], x [batch] -> [batch][n].

    Restated with the reference's masks and sums: a query outside every segment gives 0, the segment
    slope is diff(Y) ./ diff(X .- x) (the subtraction of x happens BEFORE the difference, in Float32)."""
    xr = x[:, None]                                  # x_row
    d = X - xr                                       # :73
    dYdX = np.diff(Y, axis=1) / np.diff(d, axis=1)[:, :, None]   # :74
    left = X[:, :-1]
    r = X[:, 1:]
    final_step = (r == r[:, -1:]) & (r[:, -1:] == xr)            # :79  chained .== is an elementwise AND
    mask = ((left <= xr) & (xr < r)) | final_step                  # :80
    x0 = np.zeros(X.shape[0], dtype=Y.dtype)
    y0 = np.zeros((Y.shape[0], Y.shape[2]), dtype=Y.dtype)
    dydx = np.zeros_like(y0)
    for k in range(mask.shape[1]):                                  # sums over the sequence axis, in order (:82-84)
        sel = mask[:, k]
        x0 = np.where(sel, x0 + left[:, k], x0).astype(Y.dtype)
        y0 = np.where(sel[:, None], y0 + Y[:, k, :], y0).astype(Y.dtype)
        dydx = np.where(sel[:, None], dydx + dYdX[:, k, :], dydx).astype(Y.dtype)
    return (y0 + (xr - x0[:, None]) * dydx).astype(Y.dtype)         # :86


def source_factors(t: np.ndarray, freq, dtype=F32) -> np.ndarray:
    """sin.(2.0f0 * pi * permutedims(t) * freq), src/sources.jl:22 -> [batch]."""
    return np.array([source_factor(tb, freq, dtype) for tb in t], dtype=dtype)


@dataclass
class LatentTheta:
    """θ = [C, F, PML] of get_parameters_and_initial_condition, src/model/acoustic_energy_model.jl:86-94."""
    X: np.ndarray      # [batch][nseq]      knots of C = LinearInterpolation(X, Y)
    Y: np.ndarray      # [batch][nseq][n]
    shape: np.ndarray  # [batch][n]         F = Source(shape, freq)
    freq: np.floating
    pml: np.ndarray    # [batch][n]


@dataclass
class LatentDynamics:
    """AcousticDynamics{OneDim}, src/dynamics.jl:130-149."""
    x: np.ndarray
    c0: np.floating
    grad: Gradient
    pml: np.ndarray    # build_pml(::OneDim): only pml[0] is used (dyn.pml[[1]], :192)
    bc: np.ndarray

    @staticmethod
    def make(dim: OneDim, c0, pml_width, pml_scale) -> "LatentDynamics":
        T = dim.x.dtype.type
        return LatentDynamics(dim.x, T(c0), build_gradient(dim.x), build_pml_1d(dim.x, pml_width, pml_scale),
                              build_dirichlet_1d(len(dim.x), dim.x.dtype))

    def __call__(self, w: np.ndarray, t: np.ndarray, th: LatentTheta) -> np.ndarray:
        """src/dynamics.jl:190-222.  w [batch][4][n], t [batch] -> dw [batch][4][n]."""
        T = w.dtype.type
        sigma = self.pml[0] * th.pml                                  # :192-193
        U_tot, V_tot, U_inc, V_inc = w[:, 0], w[:, 1], w[:, 2], w[:, 3]
        c = linear_interp(th.X, th.Y, t)                              # :201
        f = th.shape * source_factors(t, th.freq, w.dtype)[:, None]   # :202
        g = self.grad
        c0c = T(self.c0) * c
        dU_tot = c0c * apply_gradient(g, V_tot, -1) - sigma * U_tot                 # :210
        dV_tot = c0c * apply_gradient(g, U_tot + f, -1) - sigma * V_tot             # :211
        dU_inc = T(self.c0) * apply_gradient(g, V_inc, -1) - sigma * U_inc          # :213
        dV_inc = apply_gradient(g, U_inc + f, -1) * T(self.c0) - sigma * V_inc      # :214 (module docstring)
        return np.stack([dU_tot * self.bc, dV_tot, dU_inc * self.bc, dV_inc], axis=1).astype(w.dtype)  # :216-221


def runge_kutta(f, u, t, theta, dt):
    """src/dynamics.jl:9-16 with a vector of times (one per batch element)."""
    T = u.dtype.type
    dt = T(dt)
    hdt = T(T(0.5) * dt)
    k1 = f(u, t, theta)
    k2 = f(u + hdt * k1, (t + hdt).astype(t.dtype), theta)
    k3 = f(u + hdt * k2, (t + hdt).astype(t.dtype), theta)
    k4 = f(u + dt * k3, (t + dt).astype(t.dtype), theta)
    sixth = T(T(1.0) / T(6.0))
    du = sixth * (((k1 + T(2) * k2) + T(2) * k3) + k4)
    return du * dt


def integrate(dyn: LatentDynamics, z0: np.ndarray, tspan: np.ndarray, theta: LatentTheta, dt) -> np.ndarray:
    """(iter::Integrator)(ui, tspan::Matrix, θ), src/dynamics.jl:37-49: [time][batch][4][n]."""
    u = z0.copy()
    out = [u.copy()]
    for i in range(tspan.shape[1] - 1):
        u = (u + runge_kutta(dyn, u, np.ascontiguousarray(tspan[:, i]), theta, dt)).astype(z0.dtype)
        out.append(u.copy())
    return np.stack(out, axis=0)


def compute_latent_energy(z: np.ndarray, dx) -> np.ndarray:
    """src/model/acoustic_energy_model.jl:6-15: [batch][3][time].  Julia's sum over the first dimension has an
    unspecified (SIMD) order; the sums are accumulated in Float64 here and rounded once."""
    tot = z[:, :, 0, :].astype(np.float64)
    inc = z[:, :, 2, :].astype(np.float64)
    sc = (z[:, :, 0, :] - z[:, :, 2, :]).astype(np.float64)        # tot .- inc in Float32 first
    dx = np.float64(dx)
    e = np.stack([np.sum(tot * tot, -1), np.sum(inc * inc, -1), np.sum(sc * sc, -1)], axis=0) * dx   # [3][time][batch]
    return np.ascontiguousarray(np.transpose(e, (2, 0, 1))).astype(z.dtype)
