"""Host-side mirror of the reference's geometry / design / source types (Julia structs -> Python).

Only bookkeeping lives here (grids, cylinders, interpolators); every field computation runs in the
CUDA library.  Array convention: the reference's column-major (nx, ny, k) is held as C-order
(k, ny, nx) -- the same memory image, x fastest.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

from . import _lib

F32 = np.float32
WATER = F32(1531.0)   # src/designs.jl:13
AIR = F32(344.0)      # src/designs.jl:12


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_lib.fp)


# ---- Base.range for Float32 endpoints (rational endpoints, Float64 evaluation, one rounding) ----
def _rat(x: np.float32):
    m, y = 2048, F32(x)
    a, d, b, c = 1, 1, 0, 0
    while max(abs(a), abs(b)) <= m:
        f = int(np.trunc(y))
        y = F32(y - F32(f))
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > m:
            return c, d
        if b != 0 and F32(F32(a) / F32(b)) == F32(x):
            break
        if y == 0:
            break
        y = F32(F32(1.0) / y)
    return a, b


def julia_range(start, stop, n: int) -> np.ndarray:
    """collect(range(start, stop, n)) for Float32 (used by TwoDim src/dims.jl:56-60, build_tspan src/dynamics.jl:5-7)."""
    start, stop, n = F32(start), F32(stop), int(n)
    if n == 1 or start == stop:
        return np.full(n, start, dtype=F32)
    (sn, sd), (en, ed) = _rat(start), _rat(stop)
    if sd != 0 and ed != 0:
        den = sd * ed // math.gcd(sd, ed)
        if den != 0 and abs(den * float(start)) <= 2 ** 24 and abs(den * float(stop)) <= 2 ** 24:
            s_n, e_n = int(round(den * float(start))), int(round(den * float(stop)))
            if F32(s_n / den) == start and F32(e_n / den) == stop:
                out = np.array([float(Fraction(s_n * (n - 1 - i) + e_n * i, den * (n - 1))) for i in range(n)]).astype(F32)
                out[0], out[-1] = start, stop
                return out
    out = np.empty(n, dtype=F32)
    _lib.check(_lib.lib().waves_range_f32(C.c_float(start), C.c_float(stop), n, _fp(out)))
    return out


@dataclass
class TwoDim:
    """src/dims.jl:14-17; TwoDim(grid_size, n) :56-60."""
    x: np.ndarray
    y: np.ndarray

    def __init__(self, *args):
        if len(args) == 2 and np.ndim(args[0]) == 0:
            gs, n = F32(args[0]), int(args[1])
            self.x, self.y = julia_range(-gs, gs, n), julia_range(-gs, gs, n)
        else:
            self.x, self.y = np.ascontiguousarray(args[0], F32), np.ascontiguousarray(args[1], F32)

    def size(self):
        return (len(self.x), len(self.y))


def build_grid(dim: TwoDim) -> np.ndarray:
    """src/dims.jl:92-97 -> (2, ny, nx)."""
    nx, ny = dim.size()
    g = np.empty((2, ny, nx), dtype=F32)
    g[0], g[1] = dim.x[None, :], dim.y[:, None]
    return g


def build_wave(dim: TwoDim, fields: int = 12) -> np.ndarray:
    """src/dims.jl:107-109."""
    nx, ny = dim.size()
    return np.zeros((fields, ny, nx), dtype=F32)


def get_dx(dim) -> np.float32:
    """src/dims.jl:126."""
    return F32(_lib.lib().waves_mean_diff(_fp(dim.x), len(dim.x)))


def get_dy(dim) -> np.float32:
    """src/dims.jl:127."""
    return F32(_lib.lib().waves_mean_diff(_fp(dim.y), len(dim.y)))


def build_pml(dim: TwoDim, width, scale) -> np.ndarray:
    """The 1-D profile of build_pml(::TwoDim) (src/pml.jl:21-29)."""
    out = np.empty(len(dim.x), dtype=F32)
    _lib.check(_lib.lib().waves_build_pml_profile(_fp(dim.x), len(dim.x), C.c_float(width), C.c_float(scale), _fp(out)))
    return out


def build_gradient(dim) -> np.ndarray:
    """The distinct rows first(3) central(2) last(3) of gradient(dim.x) (src/operators.jl:10-26)."""
    out = np.empty(8, dtype=F32)
    _lib.check(_lib.lib().waves_build_gradient8(_fp(dim.x), len(dim.x), _fp(out)))
    return out


def build_normal(dim: TwoDim, mu, sigma, a) -> np.ndarray:
    """build_normal(build_grid(dim), mu, sigma, a) (src/utils.jl:12-18) -> (ny, nx)."""
    mu = np.ascontiguousarray(mu, F32).reshape(-1, 2)
    sigma = np.ascontiguousarray(sigma, F32).reshape(-1)
    a = np.ascontiguousarray(a, F32).reshape(-1)
    nx, ny = dim.size()
    out = np.empty((ny, nx), dtype=F32)
    _lib.check(_lib.lib().waves_build_normal(_fp(dim.x), nx, _fp(dim.y), ny, len(sigma), _fp(mu), _fp(sigma), _fp(a), _fp(out)))
    return out


# ---- designs (src/designs.jl) ----
class Cylinders:
    """src/designs.jl:69-88: pos (n,2), r (n,), c (n,) with the vector-space operations."""

    def __init__(self, pos, r, c):
        self.pos = np.asarray(pos, F32).reshape(-1, 2)
        self.r = np.asarray(r, F32).reshape(-1)
        self.c = np.asarray(c, F32).reshape(-1)

    def __len__(self):
        return len(self.r)

    def __add__(self, o):
        if isinstance(o, Cylinders):
            return Cylinders(self.pos + o.pos, self.r + o.r, self.c + o.c)
        return Cylinders(self.pos + F32(o), self.r + F32(o), self.c + F32(o))

    def __mul__(self, o):
        if isinstance(o, Cylinders):
            return Cylinders(self.pos * o.pos, self.r * o.r, self.c * o.c)
        return Cylinders(self.pos * F32(o), self.r * F32(o), self.c * F32(o))

    __rmul__ = __mul__

    def __sub__(self, o):
        return self + o * F32(-1.0)

    def clamp(self, lo, hi):
        return Cylinders(np.clip(self.pos, lo.pos, hi.pos), np.clip(self.r, lo.r, hi.r), np.clip(self.c, lo.c, hi.c))

    def stacked(self) -> "Cylinders":
        return self

    def table(self) -> np.ndarray:
        """(n,4) rows {x, y, r, c}: the layout waves_set_design takes."""
        return np.ascontiguousarray(np.concatenate([self.pos, self.r[:, None], self.c[:, None]], axis=1), F32)


class Cloak:
    """src/designs.jl:210-228: adjustable scatterers (`config`) around a fixed `core`."""

    def __init__(self, config: Cylinders, core: Cylinders):
        self.config, self.core = config, core

    def __add__(self, o):
        if isinstance(o, Cloak):
            return Cloak(self.config + o.config, self.core + o.core)
        if isinstance(o, Cylinders):           # Cloak + action (src/designs.jl:218)
            return Cloak(self.config + o, self.core)
        return Cloak(self.config + o, self.core + o)

    def __mul__(self, o):
        if isinstance(o, Cloak):
            return Cloak(self.config * o.config, self.core * o.core)
        return Cloak(self.config * o, self.core * o)

    __rmul__ = __mul__

    def __sub__(self, o):
        return self + o * F32(-1.0)

    def clamp(self, lo, hi):
        return Cloak(self.config.clamp(lo.config, hi.config), self.core.clamp(lo.core, hi.core))

    def stacked(self) -> Cylinders:
        """stack(config.cylinders, core) (src/designs.jl:133-138, :228)."""
        return Cylinders(np.vstack([self.config.pos, self.core.pos]), np.concatenate([self.config.r, self.core.r]),
                         np.concatenate([self.config.c, self.core.c]))

    def table(self) -> np.ndarray:
        return self.stacked().table()


class DesignSpace:
    """src/designs.jl:23-33."""

    def __init__(self, low, high):
        self.low, self.high = low, high

    def __call__(self, design, action):
        return (design + action).clamp(self.low, self.high)

    def rand(self, rng: np.random.Generator):
        """rand(space) (src/designs.jl:243-269)."""
        def cyl(lo, hi):
            u = [rng.random(a.shape, dtype=np.float32) for a in (lo.pos, lo.r, lo.c)]
            return Cylinders(u[0] * (hi.pos - lo.pos) + lo.pos, u[1] * (hi.r - lo.r) + lo.r, u[2] * (hi.c - lo.c) + lo.c)
        if isinstance(self.low, Cloak):
            return Cloak(cyl(self.low.config, self.high.config), cyl(self.low.core, self.high.core))
        return cyl(self.low, self.high)


def build_action_space(design, scale) -> DesignSpace:
    """Radii-only action box (src/designs.jl:186-191, :226)."""
    cy = design.config if isinstance(design, Cloak) else design
    s = F32(scale)
    z2, z1, one = np.zeros_like(cy.pos), np.zeros_like(cy.c), np.ones_like(cy.r)
    return DesignSpace(Cylinders(z2, one * -s, z1), Cylinders(z2.copy(), one * s, z1.copy()))


@dataclass
class DesignInterpolator:
    """src/designs.jl:274-292.  The interpolation itself runs on the device at every RK stage time."""
    initial: object
    final: object
    ti: np.float32
    tf: np.float32

    def __call__(self, t):
        ti, tf = F32(self.ti), F32(self.tf)
        dt = F32(tf - ti)
        dt = dt if dt > 0 else F32(1.0)
        s = F32(min(max(F32(t), ti), tf) - ti)
        return self.initial + ((self.final - self.initial) * F32(F32(1.0) / dt)) * s


def hexagon_ring(r) -> np.ndarray:
    """src/designs.jl:303-311."""
    r = float(F32(r))
    return np.array([[r * math.cos(i * 2 * math.pi / 6.0), r * math.sin(i * 2 * math.pi / 6.0)] for i in range(6)]).astype(F32)


def build_radii_design_space(pos: np.ndarray) -> DesignSpace:
    """src/designs.jl:337-352."""
    speed = F32(F32(3) * AIR)
    n = len(pos)
    core = Cylinders([[5.0, 0.0]], [2.0], [speed])
    return DesignSpace(Cloak(Cylinders(pos, np.full(n, 0.2), np.full(n, speed)), core),
                       Cloak(Cylinders(pos, np.full(n, 1.0), np.full(n, speed)), core))


def build_triple_ring_design_space() -> DesignSpace:
    """src/designs.jl:354-365."""
    al = 30 * math.pi / 180.0
    rot = np.array([[math.cos(al), -math.sin(al)], [math.sin(al), math.cos(al)]]).astype(F32)
    mid = hexagon_ring(4.75)
    mid = np.stack([mid[:, 0] * rot[0, 0] + mid[:, 1] * rot[1, 0], mid[:, 0] * rot[0, 1] + mid[:, 1] * rot[1, 1]], 1).astype(F32)
    rings = np.vstack([hexagon_ring(3.5), mid, hexagon_ring(6.0)])
    return build_radii_design_space((rings + np.array([[5.0, 0.0]], F32)).astype(F32))


# ---- sources (src/sources.jl) ----
class NoSource:
    """src/sources.jl:7-8."""
    shape = None
    freq = F32(0.0)

    def reset(self, rng=None):
        pass


class Source:
    """src/sources.jl:10-23."""

    def __init__(self, shape, freq):
        self.shape, self.freq = np.ascontiguousarray(shape, F32), F32(freq)

    def reset(self, rng=None):
        pass


class RandomPosGaussianSource(Source):
    """src/sources.jl:25-69: Gaussian whose centre is redrawn uniformly in [mu_low, mu_high] on reset!."""

    def __init__(self, dim: TwoDim, mu_low, mu_high, sigma, a, freq, rng=None):
        self.dim = dim
        self.mu_low, self.mu_high = np.asarray(mu_low, F32).reshape(-1, 2), np.asarray(mu_high, F32).reshape(-1, 2)
        self.sigma, self.a = np.asarray(sigma, F32), np.asarray(a, F32)
        super().__init__(build_normal(dim, self.mu_high, self.sigma, self.a), freq)
        self.reset(rng)

    def reset(self, rng=None):
        rng = rng or np.random.default_rng()
        eps = rng.random(self.mu_low.shape, dtype=np.float32)
        self.mu = (self.mu_high - self.mu_low) * eps + self.mu_low
        self.shape = build_normal(self.dim, self.mu, self.sigma, self.a)


def build_tspan(ti, dt, steps: int) -> np.ndarray:
    """src/dynamics.jl:5-7."""
    ti, dt = F32(ti), F32(dt)
    return julia_range(ti, F32(ti + F32(F32(steps) * dt)), steps + 1)
