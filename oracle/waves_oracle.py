"""CPU oracle for the Waves.jl 2-D acoustic RK4 hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement, op for op and in IEEE float32, of the
reference's CPU dynamics.  It is the *checker* for the CUDA path: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
legs may import it.  The product (`waves.jl_b200/`) never imports it and has no
CPU fallback.

PARITY UNPINNED: the reference (Julia) cannot run in this image and ships no
golden vectors for the 2-D path (its only test, test/operators.jl:4-30, pins
the 1-D gradient to within `dx` of the analytic derivative; it is ported in
tests/test_oracle.py).  Correctness of this restatement is argued from that
test, the independent RK4 statement in test/pinn.jl:38-44, and the physics
property tests in tests/.

Array convention: the reference is column-major `(nx, ny, field)`, x fastest.
Here arrays are C-ordered `(field, ny, nx)` -- the SAME memory image -- so
reference index `[i, j, k]` (1-based) is `w[k-1, j-1, i-1]`.  "along x" is the
last axis, "along y" is axis -2.

Every function cites the reference file:line (relative to /root/reference) it
restates.  All arithmetic is float32 with the reference's evaluation order
(Julia does not contract a*b+c into FMA and evaluates `a .+ b .+ c` left to
right); set `dtype=np.float64` to get the error-budget variant.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

F32 = np.float32

# src/designs.jl:8-13
ALUMINIUM = F32(3100.0)
COPPER = F32(2260.0)
BRASS = F32(2120.0)
AIR = F32(344.0)
WATER = F32(1531.0)

FIELDS = 6  # U, Vx, Vy, Psix, Psiy, Omega  (src/dynamics.jl:152-157)


# --------------------------------------------------------------------------
# Julia Base.range(start::Float32, stop::Float32, length) restated
# (Base twiceprecision.jl: rat() + exact-rational linspace, evaluated in
#  Float64 for Float32 ranges, each element rounded once to Float32)
# --------------------------------------------------------------------------
def _julia_rat_f32(x: np.float32):
    """Base.rat for a Float32 (continued fraction, |a|,|b| <= maxintfloat(Float16)=2048)."""
    m = 2048
    y = F32(x)
    a, d = 1, 1
    b, c = 0, 0
    while max(abs(a), abs(b)) <= m:
        f = int(np.trunc(y))
        y = F32(y - F32(f))
        a, c = f * a + c, a
        b, d = f * b + d, b
        if not max(abs(a), abs(b)) <= m:
            return c, d
        if b != 0 and F32(F32(a) / F32(b)) == F32(x):
            break
        if y == 0:
            break
        y = F32(F32(1.0) / y)
    return a, b


def julia_range_f32(start, stop, n: int) -> np.ndarray:
    """`collect(range(start, stop, n))` for Float32 endpoints.

    Used by TwoDim (src/dims.jl:56-60) and build_tspan (src/dynamics.jl:5-7).
    Julia finds exact rationals for the endpoints when it can and evaluates
    ref + (i-offset)*step in Float64, so each element is the Float32 rounding of
    the (near-)exact rational value.  Restated with exact Fractions; the
    non-rational fallback uses Float64 linear interpolation.
    """
    start = F32(start)
    stop = F32(stop)
    n = int(n)
    if n == 1:
        return np.array([start], dtype=F32)
    if start == stop:
        return np.full(n, start, dtype=F32)
    sn, sd = _julia_rat_f32(start)
    en, ed = _julia_rat_f32(stop)
    if sd != 0 and ed != 0:
        den = sd * ed // math.gcd(sd, ed)
        m = 2 ** 24
        if den != 0 and abs(den * float(start)) <= m and abs(den * float(stop)) <= m:
            s_n = int(round(den * float(start)))
            e_n = int(round(den * float(stop)))
            if F32(s_n / den) == start and F32(e_n / den) == stop:
                out = np.empty(n, dtype=F32)
                for i in range(n):
                    v = Fraction(s_n * (n - 1 - i) + e_n * i, den * (n - 1))
                    out[i] = F32(float(v))
                out[0] = start
                out[-1] = stop
                return out
    a = float(start)
    b = float(stop)
    out = np.array([a + i * ((b - a) / (n - 1)) for i in range(n)], dtype=np.float64).astype(F32)
    out[0] = start
    out[-1] = stop
    return out


# --------------------------------------------------------------------------
# src/dims.jl
# --------------------------------------------------------------------------
@dataclass
class TwoDim:
    """src/dims.jl:14-17, ctor :56-60."""
    x: np.ndarray
    y: np.ndarray

    @staticmethod
    def make(grid_size, n: int) -> "TwoDim":
        gs = F32(grid_size)
        return TwoDim(julia_range_f32(-gs, gs, n), julia_range_f32(-gs, gs, n))

    @property
    def shape(self):  # size(dim) = (nx, ny)
        return (len(self.x), len(self.y))


@dataclass
class OneDim:
    """src/dims.jl:8-10, ctor :48-50."""
    x: np.ndarray

    @staticmethod
    def make(grid_size, n: int) -> "OneDim":
        gs = F32(grid_size)
        return OneDim(julia_range_f32(-gs, gs, n))


def build_grid(dim: TwoDim) -> np.ndarray:
    """src/dims.jl:92-97 -> (2, ny, nx): plane 0 = x coordinate, plane 1 = y."""
    nx, ny = dim.shape
    g = np.empty((2, ny, nx), dtype=dim.x.dtype)
    g[0] = dim.x[None, :]
    g[1] = dim.y[:, None]
    return g


def build_wave(dim: TwoDim, fields: int) -> np.ndarray:
    """src/dims.jl:107-109."""
    nx, ny = dim.shape
    return np.zeros((fields, ny, nx), dtype=F32)


def build_dirichlet(dim: TwoDim) -> np.ndarray:
    """src/dims.jl:117-124: ones with a zero border -> (ny, nx)."""
    nx, ny = dim.shape
    bc = np.ones((ny, nx), dtype=dim.x.dtype)
    bc[0, :] = 0
    bc[-1, :] = 0
    bc[:, 0] = 0
    bc[:, -1] = 0
    return bc


def _mean_diff(x: np.ndarray):
    """Flux.mean(diff(x)) (src/dims.jl:126-127): Float32 pairwise sum / n.

    The pairwise order of Julia's `sum` is not restated (np.sum is also
    pairwise but blocks differently); the value agrees to ~1 ulp.
    """
    d = np.diff(x)
    return (np.sum(d, dtype=x.dtype) / x.dtype.type(len(d))).astype(x.dtype)


def get_dx(dim):
    return _mean_diff(dim.x)


def get_dy(dim):
    return _mean_diff(dim.y)


# --------------------------------------------------------------------------
# src/operators.jl
# --------------------------------------------------------------------------
@dataclass
class Gradient:
    """The three distinct rows of the sparse matrix built by
    gradient(x) (src/operators.jl:10-22): `sparse((grad / (2Δ))')`.

    first row  = [-3, 4, -1]/(2Δ)  on columns 1,2,3
    interior i = [-1, 1]/(2Δ)      on columns i-1, i+1
    last row   = [1, -4, 3]/(2Δ)   on columns n-2, n-1, n
    Each coefficient is ONE float32 division (coef / (2Δ)).
    """
    n: int
    first: np.ndarray   # 3 coefficients
    central: np.ndarray  # 2 coefficients (-k, +k)
    last: np.ndarray    # 3 coefficients
    delta: np.floating

    def to_scipy(self):
        import scipy.sparse as sp
        n = self.n
        rows, cols, vals = [0, 0, 0], [0, 1, 2], list(self.first)
        for i in range(1, n - 1):
            rows += [i, i]
            cols += [i - 1, i + 1]
            vals += list(self.central)
        rows += [n - 1] * 3
        cols += [n - 3, n - 2, n - 1]
        vals += list(self.last)
        return sp.csr_matrix((np.array(vals, dtype=self.first.dtype), (rows, cols)), shape=(n, n))


def build_gradient(x: np.ndarray) -> Gradient:
    """src/operators.jl:10-26."""
    T = x.dtype.type
    delta = T((x[-1] - x[0]) / T(len(x) - 1))
    two_delta = T(T(2) * delta)
    first = np.array([-3.0, 4.0, -1.0], dtype=x.dtype) / two_delta   # FORWARD_DIFF_COEF  :3
    last = np.array([1.0, -4.0, 3.0], dtype=x.dtype) / two_delta     # BACKWARD_DIFF_COEF :4
    central = np.array([-1.0, 1.0], dtype=x.dtype) / two_delta       # CENTRAL_DIFF_COEF  :5
    return Gradient(len(x), first, central, last, delta)


def apply_gradient(g: Gradient, u: np.ndarray, axis: int) -> np.ndarray:
    """`∇ * u` (src/operators.jl:45) along `axis`.

    SparseMatrixCSC * dense accumulates, for every output row, the products
    nzval*u in increasing COLUMN order starting from 0 (SparseArrays spmm:
    `C[row,k] += nzv * B[col,k]`, outer loop over columns of ∇), without FMA:
      interior: (c-*u[i-1]) + (c+*u[i+1])
      first:    ((c0*u[0]) + (c1*u[1])) + (c2*u[2])
      last:     ((c0*u[n-3]) + (c1*u[n-2])) + (c2*u[n-1])
    """
    u = np.moveaxis(u, axis, -1)
    out = np.empty_like(u)
    cm, cp = g.central
    out[..., 1:-1] = (cm * u[..., :-2]) + (cp * u[..., 2:])
    f0, f1, f2 = g.first
    out[..., 0] = ((f0 * u[..., 0]) + (f1 * u[..., 1])) + (f2 * u[..., 2])
    l0, l1, l2 = g.last
    out[..., -1] = ((l0 * u[..., -3]) + (l1 * u[..., -2])) + (l2 * u[..., -1])
    return np.moveaxis(out, -1, axis)


def ddx(g: Gradient, u: np.ndarray) -> np.ndarray:
    """∂x(∇,u) = ∇*u (src/operators.jl:45): along the contiguous axis (last here)."""
    return apply_gradient(g, u, -1)


def ddy(g: Gradient, u: np.ndarray) -> np.ndarray:
    """∂y(∇,u) = (∇*u')' (src/operators.jl:46): along dim 2 (axis -2 here)."""
    return apply_gradient(g, u, -2)


# --------------------------------------------------------------------------
# src/pml.jl:21-29
# --------------------------------------------------------------------------
def build_pml_profile(x: np.ndarray, width, scale) -> np.ndarray:
    """The 1-D profile of build_pml(::TwoDim): σx(i,j)=σ[i]; σy(i,j)=σ[j]
    because the dynamics uses pml' (src/dynamics.jl:161-162)."""
    T = x.dtype.type
    width = T(width)
    scale = T(scale)
    ax = np.abs(x).astype(x.dtype)
    start = T(ax[0] - width)
    region = ax > start
    out = ax.copy()
    out[~region] = 0
    if region.any():
        mn = out[region].min()
        out[region] = (out[region] - mn) / width
    # x .^ 3 * scale : Float32 literal-power 3 lowers to x*x*x
    cube = (out * out) * out
    return (cube * scale).astype(x.dtype)


def build_pml(dim: TwoDim, width, scale) -> np.ndarray:
    """src/pml.jl:21-29 -> (ny, nx) plane of σx (constant along y)."""
    prof = build_pml_profile(dim.x, width, scale)
    return np.repeat(prof[None, :], len(dim.y), axis=0)


# --------------------------------------------------------------------------
# src/utils.jl:12-18  build_normal (2-D)
# --------------------------------------------------------------------------
def _exp_like_julia(a: np.ndarray) -> np.ndarray:
    # Julia's exp(::Float32) is <1 ulp; restate as the rounding of the f64 value.
    if a.dtype == np.float64:
        return np.exp(a)
    return np.exp(a.astype(np.float64)).astype(F32)


def build_normal(grid: np.ndarray, mu: np.ndarray, sigma: np.ndarray, a: np.ndarray) -> np.ndarray:
    """src/utils.jl:12-18.  grid (2,ny,nx); mu (n,2); sigma (n,); a (n,)."""
    T = grid.dtype.type
    mu = np.asarray(mu, dtype=grid.dtype).reshape(-1, 2)
    sigma = np.asarray(sigma, dtype=grid.dtype).reshape(-1)
    a = np.asarray(a, dtype=grid.dtype).reshape(-1)
    two_pi = T(T(2.0) * T(np.pi))
    out = None
    for k in range(mu.shape[0]):
        s2 = T(sigma[k] * sigma[k])
        r2 = ((grid[0] - mu[k, 0]) ** 2) + ((grid[1] - mu[k, 1]) ** 2)   # sum(dims=3): x term then y term
        coef = T(T(1.0) / T(two_pi * s2))
        e = _exp_like_julia((-r2) / T(T(2.0) * s2))
        fk = (coef * a[k]) * e
        out = fk if out is None else out + fk
    return out.astype(grid.dtype)


# --------------------------------------------------------------------------
# src/designs.jl
# --------------------------------------------------------------------------
@dataclass
class Cylinders:
    """src/designs.jl:69-73.  pos (n,2), r (n,), c (n,)."""
    pos: np.ndarray
    r: np.ndarray
    c: np.ndarray

    def __post_init__(self):
        self.pos = np.asarray(self.pos, dtype=F32).reshape(-1, 2)
        self.r = np.asarray(self.r, dtype=F32).reshape(-1)
        self.c = np.asarray(self.c, dtype=F32).reshape(-1)

    # vector-space ops, src/designs.jl:80-88
    def __add__(self, o):
        if isinstance(o, Cylinders):
            return Cylinders(self.pos + o.pos, self.r + o.r, self.c + o.c)
        o = F32(o)
        return Cylinders(self.pos + o, self.r + o, self.c + o)

    def __mul__(self, o):
        if isinstance(o, Cylinders):
            return Cylinders(self.pos * o.pos, self.r * o.r, self.c * o.c)
        o = F32(o)
        return Cylinders(self.pos * o, self.r * o, self.c * o)

    __rmul__ = __mul__

    def __sub__(self, o):  # d1 + (-1.0f0 * d2)   src/designs.jl:48
        return self + (o * F32(-1.0))

    def __len__(self):
        return len(self.r)

    def clamp(self, low, high):
        return Cylinders(np.clip(self.pos, low.pos, high.pos), np.clip(self.r, low.r, high.r),
                         np.clip(self.c, low.c, high.c))

    def zero(self):
        return self * F32(0.0)


def stack(c1: Cylinders, c2: Cylinders) -> Cylinders:
    """src/designs.jl:133-138."""
    return Cylinders(np.vstack([c1.pos, c2.pos]), np.concatenate([c1.r, c2.r]), np.concatenate([c1.c, c2.c]))


def location_mask(cyls: Cylinders, grid: np.ndarray) -> np.ndarray:
    """src/designs.jl:99-104: strict `<` on (x-px)^2+(y-py)^2 vs r^2 -> (n,ny,nx) bool."""
    T = grid.dtype.type
    masks = np.empty((len(cyls),) + grid.shape[1:], dtype=bool)
    for k in range(len(cyls)):
        px, py = T(cyls.pos[k, 0]), T(cyls.pos[k, 1])
        r2 = T(T(cyls.r[k]) * T(cyls.r[k]))
        d2 = ((grid[0] - px) ** 2) + ((grid[1] - py) ** 2)
        masks[k] = d2 < r2
    return masks


def speed(design, grid: np.ndarray, ambient_speed):
    """src/designs.jl:110-116 (Cylinders), :63 (NoDesign), :228 (Cloak)."""
    T = grid.dtype.type
    if design is None:          # NoDesign -> scalar ambient speed
        return T(ambient_speed)
    cyls = design.all_cylinders() if hasattr(design, "all_cylinders") else design
    mask = location_mask(cyls, grid)
    ambient = (mask.sum(axis=0) == 0)
    c0 = ambient.astype(grid.dtype) * T(ambient_speed)
    c_design = np.zeros(grid.shape[1:], dtype=grid.dtype)
    for k in range(len(cyls)):  # sum(dims=3) accumulates in cylinder order
        c_design = c_design + mask[k].astype(grid.dtype) * T(cyls.c[k])
    return c0 + c_design


@dataclass
class Cloak:
    """src/designs.jl:210-228: adjustable-radii scatterers + fixed core.
    `config` plays AdjustableRadiiScatterers(cylinders) (src/designs.jl:179-181)."""
    config: Cylinders
    core: Cylinders

    def all_cylinders(self) -> Cylinders:
        return stack(self.config, self.core)   # speed(::Cloak) :228

    def __add__(self, o):
        if isinstance(o, Cloak):
            return Cloak(self.config + o.config, self.core + o.core)
        if isinstance(o, Cylinders):   # Cloak + action::AbstractScatterers  :218
            return Cloak(self.config + o, self.core)
        return Cloak(self.config + o, self.core + o)

    def __mul__(self, o):
        if isinstance(o, Cloak):
            return Cloak(self.config * o.config, self.core * o.core)
        return Cloak(self.config * o, self.core * o)

    __rmul__ = __mul__

    def __sub__(self, o):
        return self + (o * F32(-1.0))

    def clamp(self, low, high):
        return Cloak(self.config.clamp(low.config, high.config), self.core.clamp(low.core, high.core))

    def zero(self):
        return Cloak(self.config.zero(), self.core.zero())


@dataclass
class DesignSpace:
    """src/designs.jl:23-33."""
    low: object
    high: object

    def __call__(self, design, action):
        return (design + action).clamp(self.low, self.high)

    def sample(self, rng: np.random.Generator):
        """rand(space) (src/designs.jl:243-269) with a NumPy generator: only
        the radii differ between low and high in the shipped design spaces."""
        lo, hi = self.low, self.high
        if isinstance(lo, Cloak):
            u = rng.random(len(lo.config.r), dtype=np.float32)
            r = u * (hi.config.r - lo.config.r) + lo.config.r
            return Cloak(Cylinders(lo.config.pos.copy(), r, lo.config.c.copy()),
                         Cylinders(lo.core.pos.copy(), lo.core.r.copy(), lo.core.c.copy()))
        u = rng.random(len(lo.r), dtype=np.float32)
        return Cylinders(lo.pos.copy(), u * (hi.r - lo.r) + lo.r, lo.c.copy())


def build_action_space(design, scale) -> DesignSpace:
    """src/designs.jl:90-94,186-191,226: radii-only action box [-scale, scale]."""
    cyls = design.config if isinstance(design, Cloak) else design
    scale = F32(scale)
    z2 = np.zeros_like(cyls.pos)
    z1 = np.zeros_like(cyls.c)
    one = np.ones_like(cyls.r)
    return DesignSpace(Cylinders(z2, one * -scale, z1), Cylinders(z2.copy(), one * scale, z1.copy()))


@dataclass
class DesignInterpolator:
    """src/designs.jl:274-292."""
    initial: object
    final: object
    ti: np.float32
    tf: np.float32

    def __call__(self, t):
        t = F32(t)
        ti, tf = F32(self.ti), F32(self.tf)
        dt = F32(tf - ti)
        dt = dt if dt > 0 else F32(1.0)
        dy = self.final - self.initial                      # final + (-1*initial)
        slope = dy * F32(F32(1.0) / dt)                     # Δy / Δt = Δy * (1/Δt)  :49
        s = F32(min(max(t, ti), tf) - ti)                   # clamp(t,ti,tf) - ti
        return self.initial + (slope * s)                   # n*design -> design*n   :45-46


def hexagon_ring(r) -> np.ndarray:
    """src/designs.jl:303-311: angles in Float64, products rounded to Float32."""
    r = float(F32(r))
    pos = np.empty((6, 2), dtype=F32)
    for i in range(6):
        ang = i * 2 * math.pi / 6.0
        pos[i, 0] = F32(r * math.cos(ang))
        pos[i, 1] = F32(r * math.sin(ang))
    return pos


def build_2d_rotation_matrix(theta) -> np.ndarray:
    """src/designs.jl:313-319 then Float32.() at :356."""
    alpha = theta * math.pi / 180.0
    return np.array([[math.cos(alpha), -math.sin(alpha)], [math.sin(alpha), math.cos(alpha)]], dtype=np.float64).astype(F32)


def build_radii_design_space(pos: np.ndarray) -> DesignSpace:
    """src/designs.jl:337-352."""
    design_speed = F32(F32(3) * AIR)
    n = pos.shape[0]
    core = Cylinders(np.array([[5.0, 0.0]], dtype=F32), np.array([2.0], dtype=F32), np.array([design_speed], dtype=F32))
    lo = Cloak(Cylinders(pos, np.full(n, 0.2, dtype=F32), np.full(n, design_speed, dtype=F32)), core)
    hi = Cloak(Cylinders(pos.copy(), np.full(n, 1.0, dtype=F32), np.full(n, design_speed, dtype=F32)), core)
    return DesignSpace(lo, hi)


def build_triple_ring_design_space() -> DesignSpace:
    """src/designs.jl:354-365: 18 cylinders on 3 hexagonal rings around a core at (5,0)."""
    rot = build_2d_rotation_matrix(30)
    mid = hexagon_ring(4.75)
    # 6x2 * 2x2 Float32 matmul; products then one add (no FMA).  BLAS may fuse in the
    # reference; positions are inputs to the hot path so this does not affect parity of the path.
    mid_rot = np.stack([mid[:, 0] * rot[0, 0] + mid[:, 1] * rot[1, 0],
                        mid[:, 0] * rot[0, 1] + mid[:, 1] * rot[1, 1]], axis=1).astype(F32)
    rings = np.vstack([hexagon_ring(3.5), mid_rot, hexagon_ring(6.0)])
    pos = (rings + np.array([[5.0, 0.0]], dtype=F32)).astype(F32)
    return build_radii_design_space(pos)


# --------------------------------------------------------------------------
# src/sources.jl
# --------------------------------------------------------------------------
def source_factor(t, freq, dtype=F32):
    """sin(2.0f0 * pi * t * freq) (src/sources.jl:18,22,68): ((2f0*π)*t)*freq in Float32,
    Julia's sin(::Float32) restated as the rounding of the Float64 sine."""
    T = np.dtype(dtype).type
    arg = T(T(T(T(2.0) * T(np.pi)) * T(t)) * T(freq))
    if T is np.float64:
        return T(math.sin(arg))
    return F32(math.sin(float(arg)))


@dataclass
class Source:
    """src/sources.jl:10-23 (Source) and :25-69 (RandomPosGaussianSource): f(t) = shape * sin(2π t freq)."""
    shape: np.ndarray   # (ny, nx)
    freq: np.float32

    def __call__(self, t):
        return self.shape * source_factor(t, self.freq, self.shape.dtype)


class NoSource:
    """src/sources.jl:7-8."""
    def __call__(self, t):
        return F32(0.0)


# --------------------------------------------------------------------------
# src/dynamics.jl
# --------------------------------------------------------------------------
def build_tspan(ti, dt, steps: int) -> np.ndarray:
    """src/dynamics.jl:5-7."""
    ti = F32(ti)
    dt = F32(dt)
    tf = F32(ti + F32(F32(steps) * dt))
    return julia_range_f32(ti, tf, steps + 1)


@dataclass
class AcousticDynamics:
    """src/dynamics.jl:130-149."""
    dim: TwoDim
    c0: np.floating
    grad: Gradient
    pml: np.ndarray      # 1-D profile (σx(i,j)=pml[i], σy(i,j)=pml[j])
    bc: np.ndarray       # (ny,nx)

    @staticmethod
    def make(dim: TwoDim, c0, pml_width, pml_scale) -> "AcousticDynamics":
        assert len(dim.x) == len(dim.y), "σy = σx' (src/dynamics.jl:162) needs a square grid"
        return AcousticDynamics(dim, dim.x.dtype.type(c0), build_gradient(dim.x),
                                build_pml_profile(dim.x, pml_width, pml_scale), build_dirichlet(dim))

    def __call__(self, w: np.ndarray, t, theta):
        """src/dynamics.jl:179-188: θ = (C, F)."""
        C, Fsrc = theta
        c = C(t)
        f = Fsrc(t)
        dtot = acoustic_dynamics(w[0:6], c, f, self.grad, self.pml, self.bc)
        dinc = acoustic_dynamics(w[6:12], self.c0, f, self.grad, self.pml, self.bc)
        return np.concatenate([dtot, dinc], axis=0)


def acoustic_dynamics(x, c, f, grad: Gradient, pml, bc):
    """src/dynamics.jl:151-177, same evaluation order."""
    U, Vx, Vy, Px, Py, Om = x[0], x[1], x[2], x[3], x[4], x[5]
    b = c * c                      # c .^ 2  (literal power 2 -> x*x)
    sx = pml[None, :]              # σx(i,j) = σ[i]
    sy = pml[:, None]              # σy = σx'

    Vxx = ddx(grad, Vx)
    Vyy = ddy(grad, Vy)
    Uf = U + f
    Ux = ddx(grad, Uf)
    Uy = ddy(grad, Uf)

    dU = ((((b * (Vxx + Vyy)) + Px) + Py) - ((sx + sy) * U)) - Om
    dVx = Ux - (sx * Vx)
    dVy = Uy - (sy * Vy)
    dPx = (b * sx) * Vyy
    dPy = (b * sy) * Vxx
    dOm = (sx * sy) * U
    return np.stack([bc * dU, dVx, dVy, dPx, dPy, dOm], axis=0).astype(x.dtype)


def runge_kutta(f, u, t, theta, dt):
    """src/dynamics.jl:9-16; returns du (already multiplied by dt)."""
    T = u.dtype.type
    dt = T(dt)
    hdt = T(T(0.5) * dt)
    t = T(t)
    k1 = f(u, t, theta)
    k2 = f(u + hdt * k1, T(t + hdt), theta)
    k3 = f(u + hdt * k2, T(t + hdt), theta)
    k4 = f(u + dt * k3, T(t + dt), theta)
    sixth = T(T(1.0) / T(6.0))
    du = sixth * (((k1 + T(2) * k2) + T(2) * k3) + k4)
    return du * dt


def integrate(dyn, ui, tspan, theta, dt, keep=True):
    """Integrator call, src/dynamics.jl:37-53: returns (steps+1, 12, ny, nx)
    (the reference's trailing time axis is our leading one) or, with
    keep=False, only the final state."""
    u = ui.copy()
    frames = [u.copy()] if keep else None
    for i in range(len(tspan) - 1):
        u = u + runge_kutta(dyn, u, tspan[i], theta, dt)
        if keep:
            frames.append(u.copy())
    return np.stack(frames, axis=0) if keep else u


def energies(sol_U_tot: np.ndarray, sol_U_inc: np.ndarray, dim) -> np.ndarray:
    """src/env.jl:104-114 -> (frames, 3) [tot, inc, sc].  Sums are accumulated
    in Float64 and rounded (Julia's Float32 `sum` order is unspecified; treat the
    reference as this value +-1e-5 relative)."""
    T = sol_U_tot.dtype.type
    dO = T(get_dx(dim) * get_dy(dim))
    u_sc = sol_U_tot - sol_U_inc
    ax = tuple(range(1, sol_U_tot.ndim))
    tot = np.sum(sol_U_tot.astype(np.float64) ** 2, axis=ax).astype(sol_U_tot.dtype) * dO
    inc = np.sum(sol_U_inc.astype(np.float64) ** 2, axis=ax).astype(sol_U_tot.dtype) * dO
    sc = np.sum(u_sc.astype(np.float64) ** 2, axis=ax).astype(sol_U_tot.dtype) * dO
    return np.stack([tot, inc, sc], axis=1)


# --------------------------------------------------------------------------
# src/env.jl
# --------------------------------------------------------------------------
FRAMESKIP = 10  # src/env.jl:90


class WaveEnv:
    """src/env.jl:14-67 (state + ctor defaults) and :91-121 (the step)."""

    def __init__(self, dim: TwoDim, design_space, source, design=None, action_speed=250.0,
                 c0=WATER, pml_width=2.0, pml_scale=20000.0, dt=1e-5, integration_steps=100, actions=10):
        self.dim = dim
        self.design_space = design_space
        self.design = design
        self.source = source
        self.dt = F32(dt)
        self.integration_steps = int(integration_steps)
        self.actions = int(actions)
        self.action_speed = F32(action_speed)
        self.dyn = AcousticDynamics.make(dim, c0, pml_width, pml_scale)
        nx, ny = dim.shape
        self.wave = np.zeros((3, 12, ny, nx), dtype=F32)   # (nx,ny,12,3) in the reference
        self.signal = np.zeros((self.integration_steps + 1, 3), dtype=F32)
        self.time_step = 0
        self.grid = build_grid(dim)

    def time(self):  # src/env.jl:69-71  Int * Float32
        return F32(F32(self.time_step) * self.dt)

    def build_tspan(self):  # src/env.jl:73-75
        return build_tspan(self.time(), self.dt, self.integration_steps)

    def is_terminated(self):  # src/env.jl:77-79
        return self.time_step >= self.actions * self.integration_steps

    def action_scale(self):  # src/env.jl:143-145
        return F32(F32(self.action_speed * self.dt) * F32(self.integration_steps))

    def __call__(self, action, keep_frames=True):
        """src/env.jl:91-121."""
        tspan = self.build_tspan()
        ti = self.time()
        if self.design is None:
            nxt, interp = None, None
            C = lambda t: self.dyn.c0
        else:
            nxt = self.design_space(self.design, action)
            interp = DesignInterpolator(self.design, nxt, ti, tspan[-1])
            C = lambda t: speed(interp(t), self.grid, self.dyn.c0)
        sol = integrate(self.dyn, self.wave[-1], tspan, (C, self.source), self.dt)
        u_tot = sol[:, 0]
        u_inc = sol[:, 6]
        self.signal = energies(u_tot, u_inc, self.dim)
        self.design = nxt
        self.wave = sol[-(2 * FRAMESKIP) - 1::FRAMESKIP].copy()   # frames 81,91,101 (1-based)
        self.time_step += self.integration_steps
        return tspan, interp, u_tot, u_inc

    def reward(self):  # src/env.jl:147-149
        return F32(np.sum(self.signal, dtype=np.float64))


def imresize_linear(w: np.ndarray, resolution) -> np.ndarray:
    """`imresize(w, resolution)` as called by RLBase.state (src/env.jl:132-137) for w (channels, ny, nx).

    THIRD PARTY, PARITY UNPINNED: the reference uses Images.jl's imresize (ImageTransformations.jl; no Manifest, so no
    version).  Its published algorithm for shrinking: `sf = size(original) ./ size(resized)` (Int/Int -> Float64), output
    pixel i samples the BSpline(Linear()) interpolant of the original at `sf*(i - 0.5) + 0.5` (1-based; pixel-centre
    alignment, no prefilter), per axis; trailing dimensions not named in `resolution` keep their size.  Weights and
    coordinates are Float64, the result is rounded once to the element type.
    """
    c, ny, nx = w.shape
    rx, ry = int(resolution[0]), int(resolution[1])

    def taps(n_in, n_out):
        pos = (n_in / n_out) * (np.arange(n_out, dtype=np.float64) + 0.5) - 0.5   # 0-based
        i0 = np.floor(pos).astype(np.int64)
        f = pos - i0
        return np.clip(i0, 0, n_in - 1), np.clip(i0 + 1, 0, n_in - 1), f

    x0, x1, fx = taps(nx, rx)
    y0, y1, fy = taps(ny, ry)
    wd = w.astype(np.float64)
    top = (1.0 - fx) * wd[:, y0][:, :, x0] + fx * wd[:, y0][:, :, x1]
    bot = (1.0 - fx) * wd[:, y1][:, :, x0] + fx * wd[:, y1][:, :, x1]
    return ((1.0 - fy)[None, :, None] * top + fy[None, :, None] * bot).astype(w.dtype)


def flatten_repeated_last_dim(x: np.ndarray) -> np.ndarray:
    """src/utils.jl:20-31 for a list of per-action (frames, k) arrays stacked
    as (actions, frames, k): keep all frames of the first action and frames
    2..end of every later one."""
    parts = [x[0]] + [xi[1:] for xi in x[1:]]
    return np.concatenate(parts, axis=0)


def prepare_data(t_list, y_list, horizon: int):
    """prepare_data (src/data.jl:35-58) for the tspans (steps+1,) and signals (steps+1, 3) of one episode: windows of `horizon`
    consecutive actions with the repeated boundary sample dropped (flatten_repeated_last_dim, src/utils.jl:20-31)."""
    t, y = [], []
    n = horizon - 1
    for i in range(len(t_list) - n):
        t.append(flatten_repeated_last_dim(np.stack(t_list[i:i + horizon])))
        y.append(flatten_repeated_last_dim(np.stack(y_list[i:i + horizon])))
    return t, y
