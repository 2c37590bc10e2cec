"""CPU tests that pin the oracle (oracle/waves_oracle.py + the C restatement).

The reference's only real test (test/operators.jl:4-30) is ported here; the rest
are golden-fixture and physics/property checks."""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import waves_oracle as wo

F32 = np.float32


# ---- port of /root/reference/test/operators.jl:4-30 -----------------------
@pytest.mark.parametrize("fn,dfn,relative", [
    (lambda x: x ** F32(2.0), lambda x: F32(2.0) * x, False),
    (np.sin, np.cos, False),
    # The reference's third @test (y = exp(x) on [-25, 25]) cannot hold as written: the central
    # difference's truncation error f3(x)*dx^2/6 is 2.8e7 at x = 25, far above dx (the reference has
    # no runtests.jl, so nothing ever runs it).  It is ported with the bound scaled by |dy/dx|.
    (np.exp, np.exp, True),
])
def test_gradient_reference_testset(fn, dfn, relative):
    dim = wo.OneDim.make(25.0, 1024)
    dx = wo.get_dx(dim)
    grad = wo.build_gradient(dim.x)
    y = fn(dim.x).astype(F32)
    num = wo.apply_gradient(grad, y, 0)
    true = dfn(dim.x).astype(F32)
    e = num - true
    assert np.all(np.abs(e) < dx * (np.maximum(np.abs(true), 1) if relative else 1))
    # the explicit banded form is the sparse matrix of src/operators.jl:10-22
    assert np.array_equal(num, (grad.to_scipy() @ y).astype(F32))


def test_gradient_rows_match_reference_layout():
    g = wo.build_gradient(wo.TwoDim.make(15.0, 700).x)
    m = g.to_scipy().toarray()
    k = F32(1.0) / (F32(2.0) * g.delta)
    assert np.allclose(m[0, :3], np.array([-3, 4, -1]) * k) and np.count_nonzero(m[0]) == 3
    assert np.allclose(m[-1, -3:], np.array([1, -4, 3]) * k) and np.count_nonzero(m[-1]) == 3
    assert m[5, 4] == -k and m[5, 6] == k and np.count_nonzero(m[5]) == 2
    assert g.central[1] == F32(11.65)  # 1/(2dx) at 700 points over 30 m


def test_julia_range():
    x = wo.julia_range_f32(-15.0, 15.0, 700)
    assert x[0] == -15 and x[-1] == 15 and len(x) == 700
    ref = np.array([-15 + 30 * i / 699 for i in range(700)], dtype=np.float64).astype(F32)
    assert np.array_equal(x, ref)
    t = wo.build_tspan(0.0, 1e-5, 100)
    assert len(t) == 101 and t[0] == 0 and abs(float(t[-1]) - 1e-3) < 1e-9
    assert np.all(np.diff(t) > 0)


def test_pml_quirks():
    """min-subtraction makes the first PML cell 0 and the peak 19237.37."""
    dim = wo.TwoDim.make(15.0, 700)
    s = wo.build_pml_profile(dim.x, 2.0, 20000.0)
    assert (s > 0).sum() == 92 and s[46] == 0 and s[45] > 0 and s[47] == 0 and s[699 - 45] > 0
    assert abs(float(s.max()) - 19237.37) < 0.05
    assert np.array_equal(s, s[::-1])
    assert np.array_equal(wo.build_pml(dim, 2.0, 20000.0)[123], s)


def test_build_normal_peak():
    dim = wo.TwoDim.make(15.0, 700)
    sh = wo.build_normal(wo.build_grid(dim), np.array([[-10.0, 0.0]]), np.array([0.3]), np.array([1.0]))
    assert abs(float(sh.max()) - 1.7684) < 0.02 and sh.dtype == F32


def test_speed_overlap_sums_and_strict_mask():
    dim = wo.TwoDim.make(2.0, 41)   # dx = 0.1
    grid = wo.build_grid(dim)
    cyl = wo.Cylinders([[0.0, 0.0], [0.3, 0.0]], [0.5, 0.5], [1000.0, 700.0])
    c = wo.speed(cyl, grid, wo.WATER)
    assert c[20, 20] == 1700.0          # overlap: speeds add (src/designs.jl:113-115)
    assert c[0, 0] == wo.WATER
    one = wo.Cylinders([[0.0, 0.0]], [0.5], [1000.0])
    c1 = wo.speed(one, grid, wo.WATER)
    # strict '<': a cell at distance exactly r stays ambient
    assert c1[20, 25] == wo.WATER and c1[20, 24] == 1000.0
    assert wo.speed(None, grid, wo.WATER) == wo.WATER


def test_triple_ring_geometry():
    ds = wo.build_triple_ring_design_space()
    cy = ds.low.all_cylinders()
    assert len(cy) == 19 and np.all(cy.c == F32(1032.0))
    d = np.linalg.norm(cy.pos[:, None] - cy.pos[None], axis=-1) + np.eye(19) * 99
    assert d.min() > 2.0 * 1.0 + 0.4  # no overlap at r_max = 1 (core r = 2 is 3.5 away)


def test_design_interpolator_endpoints():
    ds = wo.build_triple_ring_design_space()
    rng = np.random.default_rng(0)
    d0 = ds.sample(rng)
    d1 = ds(d0, wo.build_action_space(d0, 0.25).sample(rng))
    it = wo.DesignInterpolator(d0, d1, F32(0.001), F32(0.002))
    assert np.array_equal(it(F32(0.0005)).config.r, d0.config.r)
    assert np.allclose(it(F32(0.003)).config.r, d1.config.r, atol=1e-6)
    assert np.array_equal(it(F32(0.0015)).core.r, d0.core.r)
    assert np.all(d1.config.r >= F32(0.2)) and np.all(d1.config.r <= F32(1.0))


def test_rk4_independent_statement():
    """runge_kutta vs the second statement of RK4 in test/pinn.jl:38-44 on a scalar ODE."""
    f = lambda u, t, th: -u * F32(3.0)
    u = np.array([1.0], dtype=F32)
    dt = F32(1e-2)
    du = wo.runge_kutta(f, u, F32(0.0), None, dt)
    z = 3.0 * 1e-2
    exact = 1 - z + z * z / 2 - z ** 3 / 6 + z ** 4 / 24
    assert abs(float(u[0] + du[0]) - exact) < 1e-6


# ---- golden fixtures --------------------------------------------------------
def _small(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_design_96.npz"))
    dim = wo.TwoDim.make(g["grid_size"], int(g["n"]))
    dyn = wo.AcousticDynamics.make(dim, g["c0"], g["pml_width"], g["pml_scale"])
    d0 = wo.Cylinders(g["cyl0"][:, :2], g["cyl0"][:, 2], g["cyl0"][:, 3])
    d1 = wo.Cylinders(g["cyl1"][:, :2], g["cyl1"][:, 2], g["cyl1"][:, 3])
    return g, dim, dyn, d0, d1


def test_c_oracle_matches_golden_small(golden_dir):
    g, dim, dyn, d0, d1 = _small(golden_dir)
    assert np.array_equal(dim.x, g["x"]) and np.array_equal(dyn.pml, g["sigma"])
    assert np.array_equal(co.grad8(dyn.grad), g["grad8"])
    ts = g["tspan"]
    k = co.rhs(dyn, g["u0"], g["rhs_t"], d0, d1, ts[0], ts[-1], shape=g["shape"], freq=g["freq"])
    assert np.array_equal(k, g["rhs"])
    st, en, fr = co.integrate(dyn, g["u0"], ts, g["dt"], g["dOmega"], d0, d1, ts[0], ts[-1], shape=g["shape"],
                              freq=g["freq"], save_steps=[1])
    assert np.array_equal(fr[0], g["step1"]) and np.array_equal(st, g["final"])
    assert np.allclose(en, g["energy"], rtol=1e-6, atol=0)


def test_numpy_oracle_matches_golden_small(golden_dir):
    g, dim, dyn, d0, d1 = _small(golden_dir)
    ts = g["tspan"]
    interp = wo.DesignInterpolator(d0, d1, ts[0], ts[-1])
    grid = wo.build_grid(dim)
    theta = (lambda t: wo.speed(interp(t), grid, dyn.c0), wo.Source(g["shape"], g["freq"]))
    assert np.array_equal(dyn(g["u0"], g["rhs_t"], theta), g["rhs"])
    u1 = g["u0"] + wo.runge_kutta(dyn, g["u0"], ts[0], theta, g["dt"])
    assert np.array_equal(u1, g["step1"])


def test_c_oracle_matches_golden_config1(golden_dir):
    g = np.load(os.path.join(golden_dir, "config1_700.npz"))
    dim = wo.TwoDim.make(15.0, 700)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 2.0, 20000.0)
    shape = wo.build_normal(wo.build_grid(dim), np.array([[-10.0, 0.0]]), np.array([0.3]), np.array([1.0]))
    assert np.array_equal(shape[350], g["shape_row350"])
    st, en, fr = co.integrate(dyn, np.zeros((12, 700, 700), F32), g["tspan"], 1e-5, g["dOmega"], shape=shape,
                              freq=1000.0, save_steps=[1, 10, 100])
    for i, s in enumerate((1, 10, 100)):
        assert np.array_equal(fr[i][:, g["probe_j"], g["probe_i"]], g[f"probes_{s}"])
        sums = np.array([fr[i][f].astype(np.float64).sum() for f in range(12)])
        assert np.allclose(sums, g[f"sums_{s}"], rtol=1e-12, atol=1e-12)
    assert np.allclose(en, g["energy"], rtol=1e-6, atol=0)
    # no design => both wavefields see identical inputs => bitwise equal, E_sc == 0 (SURVEY section 4 iii)
    assert np.array_equal(st[:6], st[6:]) and np.all(en[:, 2] == 0)


# ---- physics / property tests ----------------------------
def _pulse(n, gs, pml_scale, steps, transpose=False):
    dim = wo.TwoDim.make(gs, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 1.0, pml_scale)
    grid = wo.build_grid(dim)
    ic = wo.build_normal(grid, np.array([[0.7, -0.4]]), np.array([0.3]), np.array([1.0]))
    if transpose:
        ic = ic.T.copy()
    u0 = np.zeros((12, n, n), F32)
    u0[0] = ic
    u0[6] = ic
    ts = wo.build_tspan(0.0, 1e-5, steps)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    return co.integrate(dyn, u0, ts, 1e-5, dO)


def test_transpose_symmetry():
    """sigma_y = sigma_x' and one gradient for both axes => transposing the IC transposes the solution
    with Vx<->Vy and Psix<->Psiy swapped (SURVEY section 4 iv).  Not bitwise: the RHS sums Vxx+Vyy and
    Psix+Psiy in a fixed order."""
    a, ea, _ = _pulse(128, 5.0, 20000.0, 60)
    b, eb, _ = _pulse(128, 5.0, 20000.0, 60, transpose=True)
    perm = [0, 2, 1, 4, 3, 5]
    bt = np.stack([b[p].T for p in perm])
    assert np.linalg.norm(a[:6] - bt) / np.linalg.norm(a[:6]) < 1e-5
    assert np.allclose(ea, eb, rtol=1e-5)


def test_free_space_conservation_and_pml_decay():
    """scripts/pml.jl setup: with pml_scale=0 the pulse energy is (nearly) kept by the reflecting walls,
    with the PML on it is absorbed (SURVEY section 4 v / appendix sanity D)."""
    _, e_free, _ = _pulse(160, 5.0, 0.0, 700)
    _, e_pml, _ = _pulse(160, 5.0, 20000.0, 700)
    assert np.all(np.isfinite(e_free)) and np.all(np.isfinite(e_pml))
    assert e_pml[-1, 0] < 0.02 * e_pml[0, 0]
    assert e_free[-1, 0] > 10 * e_pml[-1, 0]


def test_scattering_makes_sc_energy_positive():
    dim = wo.TwoDim.make(3.0, 96)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.6, 20000.0)
    grid = wo.build_grid(dim)
    shape = wo.build_normal(grid, np.array([[-1.0, 0.0]]), np.array([0.15]), np.array([1.0]))
    cyl = wo.Cylinders([[0.3, 0.0]], [0.5], [1032.0])
    ts = wo.build_tspan(0.0, 1e-5, 150)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    st, en, _ = co.integrate(dyn, np.zeros((12, 96, 96), F32), ts, 1e-5, dO, cyl, cyl, ts[0], ts[-1], shape=shape,
                             freq=1000.0)
    assert en[-1, 2] > 1e-4 * en[-1, 1] and en[0, 2] == 0
    st0, en0, _ = co.integrate(dyn, np.zeros((12, 96, 96), F32), ts, 1e-5, dO, shape=shape, freq=1000.0)
    assert np.array_equal(st0[:6], st0[6:]) and np.array_equal(st0[6:], st[6:])  # incident field ignores the design


def test_fp64_variant_error_budget():
    """fp32 oracle vs the same code in float64: the budget the 1e-4 tolerance sits on."""
    n = 64
    x64 = np.linspace(-2.0, 2.0, n)
    dim64 = wo.TwoDim(x64, x64.copy())
    dim32 = wo.TwoDim.make(2.0, n)
    out = []
    for dim in (dim32, dim64):
        dyn = wo.AcousticDynamics.make(dim, 1531.0, 0.5, 20000.0)
        grid = wo.build_grid(dim)
        shape = wo.build_normal(grid, np.array([[-0.5, 0.0]]), np.array([0.2]), np.array([1.0]))
        src = wo.Source(shape, dim.x.dtype.type(1000.0))
        ts = wo.build_tspan(0.0, 1e-5, 50).astype(dim.x.dtype)
        u = np.zeros((12, n, n), dim.x.dtype)
        out.append(wo.integrate(dyn, u, ts, (lambda t: dyn.c0, src), dim.x.dtype.type(1e-5), keep=False))
    rel = np.linalg.norm(out[0][0] - out[1][0]) / np.linalg.norm(out[1][0])
    assert out[1].dtype == np.float64 and rel < 1e-5


def test_imresize_restatement_properties():
    """The observation resize (third party, unpinned): constants and linear ramps are reproduced exactly at the sampled
    coordinates s (i - 1/2) + 1/2, channels are independent, equal sizes are the identity."""
    n, r = 70, 16
    yy, xx = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    w = np.stack([np.full((n, n), 3.5), 2.0 * xx - 1.0, 0.5 * yy + xx]).astype(np.float32)
    out = wo.imresize_linear(w, (r, r))
    pos = (n / r) * (np.arange(r) + 0.5) - 0.5
    assert np.allclose(out[0], 3.5) and np.allclose(out[1], (2.0 * pos - 1.0)[None, :], rtol=1e-6)
    assert np.allclose(out[2], 0.5 * pos[:, None] + pos[None, :], rtol=1e-6)
    assert np.array_equal(wo.imresize_linear(w, (n, n)), w)


# ---- vectors from the real reference (absent until tests/golden/make_reference_vectors.jl has been run) ------------------
def test_oracle_matches_reference_vectors(golden_dir):
    """Pins the 2-D oracle to outputs of gladisor/Waves.jl itself.  Skips while the vectors do not exist: the image has no
    Julia, so the oracle is PARITY UNPINNED until make_reference_vectors.jl is run once elsewhere."""
    f_en = os.path.join(golden_dir, "ref_config1_energy.f32")
    f_rows = os.path.join(golden_dir, "ref_config1_final_rows.f32")
    f_small = os.path.join(golden_dir, "ref_design96_final.f32")
    if not all(os.path.exists(f) for f in (f_en, f_rows, f_small)):
        pytest.skip("no vectors from the Julia reference (run tests/golden/make_reference_vectors.jl in a Waves.jl checkout)")
    rel = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64))  # noqa: E731
    # (1) config 1
    dim = wo.TwoDim.make(15.0, 700)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 2.0, 20000.0)
    shape = wo.build_normal(wo.build_grid(dim), np.array([[-10.0, 0.0]]), np.array([0.3]), np.array([1.0]))
    ts = wo.build_tspan(0.0, 1e-5, 100)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    st, en, _ = co.integrate(dyn, np.zeros((12, 700, 700), F32), ts, 1e-5, dO, shape=shape, freq=1000.0)
    ref_en = np.fromfile(f_en, F32).reshape(101, 3)
    assert rel(en[:, :2], ref_en[:, :2]) < 1e-4 and np.all(ref_en[:, 2] <= 1e-12 * ref_en[:, 0].max())
    ref_rows = np.fromfile(f_rows, F32).reshape(12, 4, 700)          # (700, 4, 12) column-major
    assert rel(st[:, [0, 23, 349, 698], :], ref_rows) < 1e-4
    # (2) moving cylinders on 96^2
    n = 96
    dim = wo.TwoDim.make(3.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.6, 20000.0)
    grid = wo.build_grid(dim)
    pos = np.array([[0.5, 0.0], [0.9, 0.3], [-0.2, -1.1]], dtype=F32)
    d0 = wo.Cylinders(pos, [0.4, 0.3, 0.5], [1032.0, 1032.0, 2120.0])
    d1 = wo.Cylinders(pos, [0.6, 0.25, 0.35], [1032.0, 1032.0, 2120.0])
    ts = wo.build_tspan(F32(3e-4), 1e-5, 40)
    shape = wo.build_normal(grid, np.array([[-1.0, 0.2]]), np.array([0.15]), np.array([1.0]))
    u0 = np.zeros((12, n, n), F32)
    u0[0] = (F32(1e-3) * np.sin(dim.x)[None, :] * np.cos(dim.x)[:, None]).astype(F32)   # 1f-3 .* sin.(x) .* cos.(x') at [j, i]
    u0[6] = u0[0]
    st, _, _ = co.integrate(dyn, u0, ts, 1e-5, F32(wo.get_dx(dim) * wo.get_dy(dim)), d0, d1, ts[0], ts[-1], shape=shape,
                            freq=1000.0)
    assert rel(st, np.fromfile(f_small, F32).reshape(12, n, n)) < 1e-4


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(np.asarray(b, np.float64)) + 1e-300))


# ---- second, structurally independent statement of the path (oracle/sparse_oracle.py): SciPy CSC matrix products and unfused
# ---- float32 broadcasts in the reference's own (nx, ny, field) layout ---------------------------------------------------------
def test_sparse_statement_small_design_matches_oracle(golden_dir):
    """96^2, three cylinders (two overlapping) with moving radii, source, PML, random initial state, 40 steps: all 12 fields, the
    right-hand side and the energy trace of the two independent statements agree to float32 accumulation-order noise."""
    from oracle import sparse_oracle as so
    g = np.load(os.path.join(golden_dir, "small_design_96.npz"))
    dyn = so.Dynamics(g["grid_size"], int(g["n"]), g["c0"], g["pml_width"], g["pml_scale"])
    assert np.array_equal(dyn.x, g["x"]) and np.array_equal(dyn.pml[:, 0], g["sigma"])
    G = dyn.grad.toarray()
    assert np.array_equal(np.concatenate([G[0, :3], G[5, [4, 6]], G[-1, -3:]]), g["grad8"])
    shape = so.build_normal(dyn.grid, np.array([[-1.0, 0.2]]), [0.15], [1.0])
    np.testing.assert_allclose(shape.T, g["shape"], rtol=2e-6, atol=1e-30)
    ts, c0, c1 = g["tspan"], g["cyl0"], g["cyl1"]

    def C(t):
        d = so.design_at(t, c0, c1, ts[0], ts[-1])
        return so.speed(d[:, :2], d[:, 2], d[:, 3], dyn.grid, dyn.c0)

    theta = (C, so.source(np.ascontiguousarray(g["shape"].T), g["freq"]))
    u0 = so.to_julia(g["u0"])
    k = so.from_julia(dyn(u0, g["rhs_t"], theta))
    assert rel(k, g["rhs"]) < 1e-6
    sol = so.integrate(dyn, u0, ts, theta, g["dt"])
    final, step1 = so.from_julia(sol[-1]), so.from_julia(sol[1])
    assert rel(step1, g["step1"]) < 1e-6 and rel(final, g["final"]) < 1e-6
    for f in range(12):
        assert rel(final[f], g["final"][f]) < 2e-6, f
    en = so.energies(sol, F32(np.mean(np.diff(dyn.x))), F32(np.mean(np.diff(dyn.y))))
    assert np.abs(en - g["energy"]).max() / g["energy"].max() < 1e-6


def test_sparse_statement_config1_matches_golden(golden_dir):
    """BASELINE config 1 (TwoDim(15, 700), Gaussian source at (-10, 0), no design, 100 RK4 steps): the sparse-matrix statement
    against the committed fixture of the stencil statement -- all 12 fields at 512 probe points after 1, 10 and 100 steps and
    the 101 x 3 energy trace, <= 1e-6."""
    from oracle import sparse_oracle as so
    g = np.load(os.path.join(golden_dir, "config1_700.npz"))
    dyn = so.Dynamics(15.0, 700, 1531.0, 2.0, 20000.0)
    assert np.array_equal(dyn.x, g["x"])
    shape = so.build_normal(dyn.grid, np.array([[-10.0, 0.0]]), [0.3], [1.0])
    np.testing.assert_allclose(shape[:, 350], g["shape_row350"], rtol=2e-6, atol=1e-30)
    theta = (lambda t: dyn.c0, so.source(shape, 1000.0))
    sol = so.integrate(dyn, np.zeros((700, 700, 12), F32), g["tspan"], theta, 1e-5)
    for s in (1, 10, 100):
        got = so.from_julia(sol[s])[:, g["probe_j"], g["probe_i"]]
        assert rel(got, g[f"probes_{s}"]) < 1e-6, s
    en = so.energies(sol, F32(np.mean(np.diff(dyn.x))), F32(np.mean(np.diff(dyn.y))))
    assert rel(en[:, 0], g["energy"][:, 0]) < 1e-6 and rel(en[:, 1], g["energy"][:, 1]) < 1e-6
    assert np.all(en[:, 2] == 0)   # no design: both wavefields see identical inputs
