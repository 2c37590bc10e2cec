"""Host-side mirror of the reference's 1-D latent dynamics call surface.

    OneDim(grid_size, n)                                   src/dims.jl:6-10, :48-50
    AcousticDynamics(latent_dim, c0, pml_width, pml_scale) src/dynamics.jl:141-149  (the OneDim method, :190-222)
    Integrator(runge_kutta, dyn, dt)(z0, t, θ)             src/dynamics.jl:18-49    θ = [C, F, PML]
    LinearInterpolation(X, Y)                              src/utils.jl:88-98
    Source(shape, freq)                                    src/sources.jl:10-23
    compute_latent_energy(z, dx)                           src/model/acoustic_energy_model.jl:6-15
    rrule(::Integrator, z0, t, θ) / adjoint_sensitivity    src/dynamics.jl:97-128

All compute happens in libwaves_b200.so (`waves_latent_*`, csrc/latent_core.cuh); there is no CPU path.
Arrays are C-ordered views of the reference's column-major ones (same memory):
state (n,4,batch) -> [batch][4][n]; solution (n,4,batch,time) -> [time][batch][4][n]; tspan (time,batch) -> [batch][time];
X (nseq,batch) -> [batch][nseq]; Y (n,nseq,batch) -> [batch][nseq][n]; energy (time,3,batch) -> [batch][3][time].
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check
from .engine import ADJ_COMPAT, ADJ_EXACT, _ptr
from .host import F32, _fp, julia_range

LATENT_AUTO, LATENT_GENERIC, LATENT_PAIR, LATENT_ADJ_R1, LATENT_SINGLE = 0, 1, 2, 4, 8

__all__ = ["LATENT_AUTO", "LATENT_GENERIC", "LATENT_PAIR", "LATENT_ADJ_R1", "LATENT_SINGLE", "OneDim", "LinearInterpolation", "LatentSource", "LatentDynamics", "LatentIntegrator", "build_pml_1d", "ADJ_EXACT",
           "ADJ_COMPAT"]


@dataclass
class OneDim:
    """src/dims.jl:6-10; OneDim(grid_size, n) :48-50."""
    x: np.ndarray

    def __init__(self, *args):
        if len(args) == 2:
            gs = F32(args[0])
            self.x = julia_range(-gs, gs, int(args[1]))
        else:
            self.x = np.ascontiguousarray(args[0], F32)

    def size(self):
        return (len(self.x),)


def build_pml_1d(dim: OneDim, width, scale) -> np.ndarray:
    """build_pml(::OneDim, width, scale), src/pml.jl:6-15."""
    out = np.empty(len(dim.x), dtype=F32)
    check(_lib.lib().waves_latent_build_pml(_fp(dim.x), len(dim.x), C.c_float(width), C.c_float(scale), _fp(out), None))
    return out


@dataclass
class LinearInterpolation:
    """θ[1] = C (src/utils.jl:88-98): X [batch][nseq] knots, Y [batch][nseq][n] values."""
    X: np.ndarray
    Y: np.ndarray


@dataclass
class LatentSource:
    """θ[2] = F = Source(shape, freq) with a (n, batch) shape (src/sources.jl:10-23); shape None: no source."""
    shape: np.ndarray | None
    freq: float


class LatentDynamics:
    """AcousticDynamics{OneDim} (src/dynamics.jl:130-149): constants of the 1-D right-hand side."""

    def __init__(self, dim: OneDim, c0, pml_width, pml_scale):
        self.dim, self.c0 = dim, F32(c0)
        self.pml_width, self.pml_scale = F32(pml_width), F32(pml_scale)
        self.pml = build_pml_1d(dim, pml_width, pml_scale)


class LatentIntegrator:
    """Integrator(runge_kutta, dyn::AcousticDynamics{OneDim}, dt) (src/dynamics.jl:18-49) on one GPU."""

    def __init__(self, dyn: LatentDynamics, dt, device=0):
        self.dyn, self.dt, self.device = dyn, F32(dt), int(device)
        self.n = len(dyn.dim.x)
        self._x = np.ascontiguousarray(dyn.dim.x, F32)
        cfg = _lib.LatentConfig(n=self.n, device=self.device, c0=float(dyn.c0), dt=float(self.dt),
                                pml_width=float(dyn.pml_width), pml_scale=float(dyn.pml_scale), pml0=-1.0, dx=0.0,
                                x=self._x.ctypes.data_as(_lib.fp), grad8=None)
        h = C.c_void_p()
        check(_lib.lib().waves_latent_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().waves_latent_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _f32(a):
        return a if (a is None or hasattr(a, "data_ptr")) else np.ascontiguousarray(a, F32)

    def _tspan(self, tspan, batch):
        """tspan [batch][time]; a vector is iter(ui, tspan::AbstractVector, θ) = iter(ui, tspan[:, :], θ) (src/dynamics.jl:51-53):
        the same times for every batch element.  Host arrays and CUDA tensors alike."""
        tspan = self._f32(tspan)
        if tspan.ndim == 1:
            if hasattr(tspan, "data_ptr"):
                return tspan[None, :].expand(batch, -1).contiguous()
            return np.ascontiguousarray(np.broadcast_to(tspan[None, :], (batch, len(tspan))))
        return tspan

    def _theta(self, theta):
        Cint, Fsrc, pml = theta
        X, Y = self._f32(Cint.X), self._f32(Cint.Y)
        shape = self._f32(Fsrc.shape) if Fsrc is not None else None
        freq = float(Fsrc.freq) if Fsrc is not None else 0.0
        return X, Y, shape, freq, self._f32(pml)

    def __call__(self, z0, tspan, theta, want_z=True, want_energy=False, out_z=None, out_energy=None):
        """iter(z0, tspan, θ) (src/dynamics.jl:37-49) -> z [time][batch][4][n]; with want_energy also
        compute_latent_energy(z, dx) [batch][3][time] (src/model/acoustic_energy_model.jl:6-15).  With want_z=False the
        trajectory never reaches HBM and only the energies (and the last state) come back."""
        z0 = self._f32(z0)
        tspan = self._tspan(tspan, int(z0.shape[0]))
        X, Y, shape, freq, pml = self._theta(theta)
        batch, steps, nseq = int(z0.shape[0]), int(tspan.shape[1]) - 1, int(X.shape[1])
        assert tuple(z0.shape) == (batch, 4, self.n) and tuple(Y.shape) == (batch, nseq, self.n)
        z = out_z if out_z is not None else (np.empty((steps + 1, batch, 4, self.n), F32) if want_z else None)
        e = out_energy if out_energy is not None else (np.empty((batch, 3, steps + 1), F32) if want_energy else None)
        last = None if want_z else np.empty((batch, 4, self.n), F32)
        check(_lib.lib().waves_latent_integrate(self._h, batch, steps, nseq, _ptr(z0), _ptr(tspan), _ptr(X), _ptr(Y),
                                                _ptr(shape), C.c_float(freq), _ptr(pml), _ptr(z), _ptr(e), _ptr(last)))
        out = z if want_z else last
        return (out, e) if (want_energy or out_energy is not None) else out

    def adjoint(self, z, tspan, theta, w_energy=None, dL_dz=None, mode=ADJ_EXACT):
        """adjoint_sensitivity(iter, z, t, θ, ∂L_∂z) (src/dynamics.jl:97-118) -> dict(z0, Y, shape, pml)."""
        z = self._f32(z)
        tspan = self._tspan(tspan, int(z.shape[1]))
        X, Y, shape, freq, pml = self._theta(theta)
        batch, steps, nseq = int(z.shape[1]), int(tspan.shape[1]) - 1, int(X.shape[1])
        we, gz = self._f32(w_energy), self._f32(dL_dz)   # keep the converted copies alive across the call
        g = {"z0": np.empty((batch, 4, self.n), F32), "Y": np.empty((batch, nseq, self.n), F32),
             "shape": np.empty((batch, self.n), F32) if shape is not None else None, "pml": np.empty((batch, self.n), F32)}
        check(_lib.lib().waves_latent_adjoint(self._h, batch, steps, nseq, _ptr(z), _ptr(tspan), _ptr(X), _ptr(Y), _ptr(shape),
                                              C.c_float(freq), _ptr(pml), int(mode), _ptr(we),
                                              _ptr(gz), _ptr(g["z0"]), _ptr(g["Y"]), _ptr(g["shape"]),
                                              _ptr(g["pml"])))
        return g

    def set_generic(self, on: bool):
        """Force the generic shared-memory kernels (the register fast paths are the default where they apply)."""
        check(_lib.lib().waves_latent_set_generic(self._h, int(bool(on))))

    def set_variant(self, variant: int):
        """Flags: LATENT_AUTO (default: the fastest kernels measured on a B200 -- the pair forms of both passes), LATENT_GENERIC,
        or an explicit choice: LATENT_SINGLE (one element per thread) / LATENT_PAIR (two) for the forward pass, | LATENT_ADJ_R1
        for the register reverse kernel (otherwise the generic one)."""
        check(_lib.lib().waves_latent_set_variant(self._h, int(variant)))

    def last_kernel_ms(self) -> float:
        """Device time of the kernel of the last call (CUDA events on the handle's stream)."""
        return float(_lib.lib().waves_latent_last_kernel_ms(self._h))

    def launch_count(self) -> int:
        return int(_lib.lib().waves_latent_launch_count(self._h))

