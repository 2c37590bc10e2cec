"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs, the committed golden fixtures, and size-independent properties at the
BASELINE sizes.  Tolerance (BASELINE.json north_star): relative L2 <= 1e-4 on fields and energy traces;
WAVES_MODE_EXACT is additionally required to be BIT-EXACT with the oracle."""
import os

import numpy as np
import pytest

import waves_b200 as wb
from oracle import c_oracle as co
from oracle import waves_oracle as wo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32
TOL = 1e-4        # stated tolerance
TIGHT = 2e-5      # what the fused kernel actually delivers (FMA contraction + reassociated stencils)


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b.astype(np.float64)) + 1e-300))


def table(cyl):
    return np.ascontiguousarray(np.concatenate([cyl.pos, cyl.r[:, None], cyl.c[:, None]], axis=1), F32)


def make_engine(dim, dyn, n_env=1, dO=None):
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim)) if dO is None else dO
    return wb.Engine(dim.x, dim.y, dyn.c0, 1e-5, n_env=n_env, sigma=dyn.pml, grad8=co.grad8(dyn.grad), d_omega=float(dO))


def load_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_design_96.npz"))
    dim = wo.TwoDim.make(g["grid_size"], int(g["n"]))
    dyn = wo.AcousticDynamics.make(dim, g["c0"], g["pml_width"], g["pml_scale"])
    return g, dim, dyn


@pytest.mark.parametrize("mode", [wb.MODE_EXACT, wb.MODE_FUSED])
def test_golden_small(golden_dir, mode):
    g, dim, dyn = load_small(golden_dir)
    eng = make_engine(dim, dyn, dO=g["dOmega"])
    ts = g["tspan"]
    eng.set_state(g["u0"][None])
    eng.set_source(g["shape"], float(g["freq"]))
    eng.set_design(g["cyl0"], g["cyl1"], ts[0], ts[-1])
    k = eng.rhs(g["rhs_t"])
    assert np.array_equal(k, g["rhs"]), "dyn(x,t,θ) must be bit-exact"
    en, fr = eng.integrate(ts, mode, energy=True, save_steps=[1, len(ts) - 1])
    final = eng.get_state(0)
    if mode == wb.MODE_EXACT:
        assert np.array_equal(fr[0, 0], g["step1"]) and np.array_equal(final, g["final"])
        assert np.allclose(en[0], g["energy"], rtol=1e-6, atol=0)
    else:
        assert rel(fr[0, 0], g["step1"]) < TIGHT and rel(final, g["final"]) < TIGHT
        for f in range(12):
            assert rel(final[f], g["final"][f]) < TOL, f"field {f}"
        assert np.abs(en[0] - g["energy"]).max() / g["energy"].max() < TIGHT
    assert np.array_equal(fr[0, 1], final)
    eng.close()


def test_golden_config1_fused(golden_dir):
    """BASELINE config 1: TwoDim(15,700), Gaussian source, no design, 100 steps."""
    g = np.load(os.path.join(golden_dir, "config1_700.npz"))
    dim = wb.TwoDim(15.0, 700)
    assert np.array_equal(dim.x, g["x"])
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, d_omega=float(g["dOmega"]))
    shape = wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0])
    assert np.array_equal(shape[350], g["shape_row350"])
    eng.set_source(shape, 1000.0)
    en, fr = eng.integrate(g["tspan"], wb.MODE_FUSED, energy=True, save_steps=[1, 10, 100])
    for i, s in enumerate((1, 10, 100)):
        got, want = fr[0, i][:, g["probe_j"], g["probe_i"]], g[f"probes_{s}"]
        assert rel(got, want) < TIGHT   # (plain field sums cancel to rounding noise: only the bit-exact oracle test pins them)
    e, ge = en[0], g["energy"]
    assert rel(e[:, 0], ge[:, 0]) < TIGHT and rel(e[:, 1], ge[:, 1]) < TIGHT
    # no design: both wavefields see identical inputs -> bitwise equal, scattered energy exactly 0
    final = fr[0, 2]
    assert np.array_equal(final[:6], final[6:]) and np.all(e[:, 2] == 0)
    eng.close()


@pytest.mark.parametrize("n,gs,pw,dt", [(33, 1.0, 0.3, 5e-6), (70, 2.0, 0.5, 1e-5), (100, 3.0, 0.0, 1e-5), (257, 6.0, 1.0, 1e-5)])
def test_ragged_sizes_vs_oracle(n, gs, pw, dt):
    """Edge cases: nx not a multiple of 4 / of the 24-column strip, tiny grids, pml_width = 0."""
    dim = wo.TwoDim.make(gs, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, pw if pw > 0 else 0.5, 20000.0 if pw > 0 else 0.0)
    rng = np.random.default_rng(n)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(F32)
    shape = wo.build_normal(wo.build_grid(dim), np.array([[-0.3 * gs, 0.1 * gs]]), np.array([0.08 * gs]), np.array([1.0]))
    pos = np.array([[0.2 * gs, 0.0], [0.35 * gs, 0.1 * gs], [-0.1 * gs, -0.4 * gs]], F32)
    d0 = wo.Cylinders(pos, np.array([0.15, 0.1, 0.2]) * gs, [1032.0, 1032.0, 2120.0])
    d1 = wo.Cylinders(pos, np.array([0.2, 0.08, 0.12]) * gs, [1032.0, 1032.0, 2120.0])
    ts = wo.build_tspan(F32(2e-4), F32(dt), 12)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    ref, ren, _ = co.integrate(dyn, u0, ts, dt, dO, d0, d1, ts[0], ts[-1], shape=shape, freq=1000.0)
    for mode in (wb.MODE_EXACT, wb.MODE_FUSED):
        eng = wb.Engine(dim.x, dim.y, dyn.c0, dt, sigma=dyn.pml, grad8=co.grad8(dyn.grad), d_omega=float(dO))
        eng.set_state(u0[None])
        eng.set_source(shape, 1000.0)
        eng.set_design(table(d0), table(d1), ts[0], ts[-1])
        en, _ = eng.integrate(ts, mode)
        got = eng.get_state(0)
        if mode == wb.MODE_EXACT:
            assert np.array_equal(got, ref)
        else:
            assert rel(got, ref) < TIGHT
            for f in range(12):
                assert rel(got[f], ref[f]) < TOL
        assert np.abs(en[0] - ren).max() / ren.max() < TIGHT
        eng.close()


def test_batch_of_independent_envs():
    """config 3 in miniature: every env has its own state, design and source; no cross-talk."""
    n, E = 96, 5
    dim = wo.TwoDim.make(3.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.6, 20000.0)
    grid = wo.build_grid(dim)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    ts = wo.build_tspan(F32(1e-4), F32(1e-5), 15)
    eng = make_engine(dim, dyn, n_env=E)
    refs = []
    for e in range(E):
        rng = np.random.default_rng(100 + e)
        u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(F32)
        shape = wo.build_normal(grid, rng.uniform(-1.5, 1.5, (1, 2)), np.array([0.15]), np.array([1.0]))
        pos = rng.uniform(-1.2, 1.2, (4, 2)).astype(F32)
        d0 = wo.Cylinders(pos, rng.uniform(0.2, 0.5, 4), [1032.0] * 4)
        d1 = wo.Cylinders(pos, rng.uniform(0.2, 0.5, 4), [1032.0] * 4)
        eng.set_state(u0[None], env=e)
        if e == 2:   # one env without source, one without design
            eng.set_source(None, 0.0, env=e)
            shape = None
        else:
            eng.set_source(shape, 900.0 + 50 * e, env=e)
        if e == 3:
            eng.set_design(None, None, 0, 0, env=e)
            d0 = d1 = None
        else:
            eng.set_design(table(d0), table(d1), ts[0], ts[-1], env=e)
        refs.append(co.integrate(dyn, u0, ts, 1e-5, dO, d0, d1, ts[0], ts[-1], shape=shape, freq=900.0 + 50 * e))
    en, _ = eng.integrate(ts, wb.MODE_FUSED)
    got = eng.get_state(-1)
    for e in range(E):
        assert rel(got[e], refs[e][0]) < TIGHT, f"env {e}"
        assert np.abs(en[e] - refs[e][1]).max() / refs[e][1].max() < TIGHT
    eng.close()


def test_waveenv_mirror_matches_oracle_env():
    """(env::WaveEnv)(action) (src/env.jl:91-121): signal, kept frames, design and time bookkeeping."""
    n = 128
    dimo, dimg = wo.TwoDim.make(15.0, n), wb.TwoDim(15.0, n)
    dso, dsg = wo.build_triple_ring_design_space(), wb.build_triple_ring_design_space()
    rng = np.random.default_rng(2)
    d_init = dsg.rand(rng)
    shape = wb.build_normal(dimg, [[-6.0, 1.0]], [0.8], [1.0])
    envg = wb.WaveEnv(dimg, design_space=dsg, source=wb.Source(shape, 1000.0), dt=2e-5, integration_steps=30, actions=3,
                      resolution=(64, 64))   # src/env.jl:52: the resolution must be below the grid size
    envg.design = d_init
    envo = wo.WaveEnv(dimo, dso, wo.Source(shape, F32(1000.0)), dt=2e-5, integration_steps=30, actions=3,
                      design=wo.Cloak(wo.Cylinders(d_init.config.pos, d_init.config.r, d_init.config.c),
                                      wo.Cylinders(d_init.core.pos, d_init.core.r, d_init.core.c)))
    sig_g, sig_o = [], []
    while not envg.is_terminated():
        act = envg.action_space().rand(rng)
        tg, _, ut, ui = envg(act, return_frames=True)
        to, _, uto, uio = envo(wo.Cylinders(act.pos, act.r, act.c))
        assert np.array_equal(tg, to)
        assert rel(ut, uto) < TIGHT and rel(ui, uio) < TIGHT
        assert rel(envg.wave, envo.wave) < TIGHT
        sig_g.append(envg.signal)
        sig_o.append(envo.signal)
    assert envo.is_terminated() and envg.time_step == envo.time_step == 90
    sg, so = wo.flatten_repeated_last_dim(np.stack(sig_g)), wo.flatten_repeated_last_dim(np.stack(sig_o))
    assert sg.shape == (91, 3) and rel(sg, so) < TIGHT
    assert so[-1, 2] > 0, "the pulse must have reached the scatterers"
    assert abs(float(envg.reward()) - float(envo.reward())) <= 1e-4 * abs(float(envo.reward()))


def test_integrator_mirror_returns_reference_shape():
    dim = wb.TwoDim(5.0, 128)
    it = wb.Integrator(wb.AcousticDynamics(dim, wb.WATER, 1.0, 0.0), 1e-5)
    wave = wb.build_wave(dim, 12)
    ic = wb.build_normal(dim, [[0.0, 0.0]], [0.3], [1.0])
    wave[0] = ic
    wave[6] = ic
    ts = it.build_tspan(0.0, 20)
    sol = it(wave, ts, (None, wb.NoSource()))       # scripts/pml.jl:6-18
    assert sol.shape == (21, 12, 128, 128) and np.array_equal(sol[0], wave)
    assert np.array_equal(sol[:, :6], sol[:, 6:])


# ---- BASELINE-size properties (700^2), independent of any oracle run --------------------------
def _engine700(n_env=1):
    dim = wb.TwoDim(15.0, 700)
    return dim, wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=n_env)


def test_full_size_linearity_and_fused_vs_exact():
    """Without a source the step is linear: scaling the state by 2 scales the result by exactly 2 (bitwise);
    and the fused kernel agrees with the exact per-stage kernels at full size."""
    dim, eng = _engine700(2)
    ds = wb.build_triple_ring_design_space()
    rng = np.random.default_rng(11)
    d0 = ds.rand(rng)
    d1 = ds(d0, wb.build_action_space(d0, 0.25).rand(rng))
    u0 = (rng.standard_normal((12, 700, 700)) * 1e-3).astype(F32)
    ts = wb.build_tspan(0.0, 1e-5, 6)
    eng.set_design(d0.table(), d1.table(), ts[0], ts[-1])
    eng.set_state(u0[None], env=0)
    eng.set_state((2 * u0)[None], env=1)
    eng.integrate(ts, wb.MODE_FUSED, energy=False)
    a = eng.get_state(-1)
    assert np.array_equal(2 * a[0], a[1])
    eng.set_state(u0[None], env=0)
    eng.set_state(u0[None], env=1)
    eng.integrate(ts, wb.MODE_EXACT, energy=False)
    b = eng.get_state(0)
    assert rel(a[0], b) < TIGHT
    eng.close()


def test_full_size_transpose_symmetry():
    dim, eng = _engine700(2)
    ic = wb.build_normal(dim, [[3.0, -5.0]], [0.4], [1.0])
    u = np.zeros((2, 12, 700, 700), F32)
    u[0, 0] = u[0, 6] = ic
    u[1, 0] = u[1, 6] = ic.T
    eng.set_state(u)
    en, _ = eng.integrate(wb.build_tspan(0.0, 1e-5, 40), wb.MODE_FUSED)
    out = eng.get_state(-1)
    perm = [0, 2, 1, 4, 3, 5]
    bt = np.stack([out[1, p].T for p in perm])
    assert rel(out[0, :6], bt) < TIGHT and np.allclose(en[0], en[1], rtol=1e-5)
    eng.close()


def test_pml_absorbs_and_walls_reflect():
    """scripts/pml.jl: free pulse, energy kept with pml_scale = 0, absorbed with the PML on."""
    dim = wb.TwoDim(5.0, 160)
    ic = wb.build_normal(dim, [[0.7, -0.4]], [0.3], [1.0])
    u = np.zeros((1, 12, 160, 160), F32)
    u[0, 0] = u[0, 6] = ic
    ts = wb.build_tspan(0.0, 1e-5, 700)
    ends = []
    for scale in (0.0, 20000.0):
        eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 1.0, scale)
        eng.set_state(u)
        en, _ = eng.integrate(ts, wb.MODE_FUSED)
        assert np.all(np.isfinite(en))
        ends.append((en[0, 0, 0], en[0, -1, 0]))
        eng.close()
    assert ends[1][1] < 0.02 * ends[1][0] and ends[0][1] > 10 * ends[1][1]


def test_speed_plane_and_error_behaviour():
    n = 64
    dim = wo.TwoDim.make(2.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.4, 20000.0)
    rng = np.random.default_rng(9)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(F32)
    cpl = (1531.0 - 400.0 * np.exp(-((wo.build_grid(dim)[0]) ** 2 + (wo.build_grid(dim)[1]) ** 2))).astype(F32)
    ts = wo.build_tspan(F32(0.0), F32(1e-5), 5)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    ref, _, _ = co.integrate(dyn, u0, ts, 1e-5, dO, cplane=cpl)
    eng = make_engine(dim, dyn)
    eng.set_state(u0[None])
    eng.set_speed_field(cpl)
    eng.integrate(ts, wb.MODE_EXACT, energy=False)
    assert np.array_equal(eng.get_state(0), ref)
    with pytest.raises(wb.WavesError, match="speed plane"):
        eng.integrate(ts, wb.MODE_FUSED, energy=False)
    with pytest.raises(wb.WavesError):
        eng.integrate(ts, 7, energy=False)
    with pytest.raises(wb.WavesError, match="ascending"):
        eng.set_speed_field(None)
        eng.integrate(ts, wb.MODE_FUSED, energy=False, save_steps=[3, 2])
    with pytest.raises(wb.WavesError, match="out of range"):
        eng.set_state(u0[None], env=4)
    eng.close()
    with pytest.raises(wb.WavesError, match="sigma_y"):
        wb.Engine(dim.x, dim.y[:-1], 1531.0, 1e-5)


def test_many_cylinders_overflow_path():
    """More cylinders under one warp window than its culled list holds (slow exact loop)."""
    n = 96
    dim = wo.TwoDim.make(3.0, n)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 0.6, 20000.0)
    rng = np.random.default_rng(21)
    pos = rng.uniform(-0.8, 0.8, (40, 2)).astype(F32)
    d0 = wo.Cylinders(pos, rng.uniform(0.1, 0.4, 40), rng.uniform(900, 1400, 40))
    d1 = wo.Cylinders(pos, rng.uniform(0.1, 0.4, 40), d0.c)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(F32)
    ts = wo.build_tspan(F32(0.0), F32(1e-5), 6)
    dO = F32(wo.get_dx(dim) * wo.get_dy(dim))
    ref, ren, _ = co.integrate(dyn, u0, ts, 1e-5, dO, d0, d1, ts[0], ts[-1])
    eng = make_engine(dim, dyn)
    eng.set_state(u0[None])
    eng.set_design(table(d0), table(d1), ts[0], ts[-1])
    en, _ = eng.integrate(ts, wb.MODE_FUSED)
    assert rel(eng.get_state(0), ref) < TIGHT
    eng.close()


def test_constant_auxiliary_fields_shortcut_is_invalidated_by_state_writes(golden_dir):
    """Where sigma is zero, Psi/Omega never change: the fused path stops copying them (and reads their sum from its own
    plane) after the first step.  Writing a new state must invalidate that: run A, then B with different auxiliary fields on
    the SAME handle, each against the exact per-stage kernels."""
    g, dim, dyn = load_small(golden_dir)
    ts = g["tspan"][:13]
    rng = np.random.default_rng(11)
    states = [g["u0"], (rng.standard_normal(g["u0"].shape) * 2e-3).astype(F32)]
    fused, exact = make_engine(dim, dyn, dO=g["dOmega"]), make_engine(dim, dyn, dO=g["dOmega"])
    for eng in (fused, exact):
        eng.set_source(g["shape"], float(g["freq"]))
        eng.set_design(g["cyl0"], g["cyl1"], ts[0], ts[-1])
    for u0 in states:
        fused.set_state(u0[None])
        exact.set_state(u0[None])
        ef, _ = fused.integrate(ts, wb.MODE_FUSED)
        ee, _ = exact.integrate(ts, wb.MODE_EXACT)
        a, b = fused.get_state(0), exact.get_state(0)
        for f in range(12):
            assert rel(a[f], b[f]) < TIGHT, f"field {f}"
        assert np.abs(ef - ee).max() / ee.max() < TIGHT
        # single steps after an integration keep working on the shortcut path too
        fused.step(float(ts[-1]))
        exact.step(float(ts[-1]), wb.MODE_EXACT)
        assert rel(fused.get_state(0), exact.get_state(0)) < TIGHT
    fused.close()
    exact.close()


def test_full_size_waveenv_actions_match_c_oracle():
    """BASELINE configs[1] at full size (700^2, triple-ring design, Gaussian source): 3 env(action) calls of 100 RK4
    steps through the WaveEnv mirror against the C restatement chained the same way (signal, kept frames, final state)."""
    n, steps = 700, 100
    dimg, dimo = wb.TwoDim(15.0, n), wo.TwoDim.make(15.0, n)
    dyn = wo.AcousticDynamics.make(dimo, wo.WATER, 2.0, 20000.0)
    dO = F32(wo.get_dx(dimo) * wo.get_dy(dimo))
    rng = np.random.default_rng(4)
    ds = wb.build_triple_ring_design_space()
    # the ring (centred at x = 5) is 8 m from the reference's source line x = -10: 300 steps would not reach it, so the
    # source sits closer here and the scattered energy is exercised
    shape = wb.build_normal(dimg, [[-3.5, 2.5]], [0.3], [1.0])
    env = wb.WaveEnv(dimg, design_space=ds, source=wb.Source(shape, 1000.0), integration_steps=steps, actions=3, rng=rng)
    as_cyl = lambda d: wo.Cylinders(d.table()[:, :2], d.table()[:, 2], d.table()[:, 3])
    u = np.zeros((12, n, n), F32)
    while not env.is_terminated():
        d0 = env.design
        act = env.action_space().rand(rng)
        ts, interp, _, _ = env(act)
        u, en, fr = co.integrate(dyn, u, ts, 1e-5, dO, as_cyl(d0), as_cyl(env.design), ts[0], ts[-1], shape=shape, freq=1000.0,
                                 save_steps=[steps - 20, steps - 10, steps])
        scale = en[:, :2].max()
        assert np.abs(env.signal - en).max() / scale < TIGHT
        assert rel(env.wave[-1], u) < TIGHT and rel(env.wave[0], fr[0]) < TIGHT
    assert env.time_step == 300 and en[-1, 2] > 0


def test_observation_image_matches_restated_imresize():
    """RLBase.state(env) (src/env.jl:132-137): device-side imresize of (3 U_tot frames + source shape) vs the NumPy restatement;
    host and device frame blocks, a batch of environments, and the WaveEnv mirror."""
    import torch
    n = 200
    dim = wb.TwoDim(6.0, n)
    rng = np.random.default_rng(9)
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 1.0, 20000.0, n_env=2)
    shapes = [wb.build_normal(dim, [[-2.0, 0.5 * e]], [0.4], [1.0]) for e in range(2)]
    for e in range(2):
        eng.set_source(shapes[e], 1000.0, env=e)
    frames = rng.standard_normal((2, 3, 12, n, n)).astype(F32)
    for res in ((128, 128), (64, 96)):
        got = eng.observe(frames, res)
        for e in range(2):
            want = wo.imresize_linear(np.concatenate([frames[e, :, 0], shapes[e][None]]), res)
            assert got[e].shape == want.shape and np.abs(got[e] - want).max() <= 1e-6 * np.abs(want).max()
        dev = eng.observe(torch.from_numpy(frames).cuda(), res, out=torch.empty((2, 4, res[1], res[0]), device="cuda"))
        assert np.array_equal(dev.cpu().numpy(), got)
    eng.close()
    env = wb.WaveEnv(dim, design_space=wb.build_triple_ring_design_space(), source=wb.Source(shapes[0], 1000.0),
                     integration_steps=20, actions=2, pml_width=1.0, rng=rng)
    env(env.action_space().rand(rng))
    ts, x, design = env.state()
    want = wo.imresize_linear(np.concatenate([env.wave[:, 0], shapes[0][None]]), (128, 128))
    assert x.shape == (4, 128, 128) and np.abs(x - want).max() <= 1e-6 * max(np.abs(want).max(), 1e-30)


def test_batch_env_and_episode_layer():
    """BatchWaveEnv: environments stepping in lockstep on one handle are bitwise the single-environment WaveEnv; the episode
    layer (generate_episode!, prepare_data: src/data.jl:12-58) on top of both."""
    n, steps, actions, E = 128, 20, 4, 3
    dim = wb.TwoDim(15.0, n)
    ds = wb.build_triple_ring_design_space()
    shapes = [wb.build_normal(dim, [[-6.0 + e, 1.0 - e]], [0.8], [1.0]) for e in range(E)]
    kw = dict(dt=2e-5, integration_steps=steps, actions=actions, resolution=(32, 32))
    benv = wb.BatchWaveEnv(dim, ds, [wb.Source(s, 1000.0) for s in shapes], rngs=[np.random.default_rng(10 + e) for e in range(E)], **kw)
    pol_rngs = [np.random.default_rng(100 + e) for e in range(E)]
    eps = wb.generate_episodes([lambda env, e=e: env.action_spaces()[e].rand(pol_rngs[e]) for e in range(E)], benv)
    assert benv.is_terminated() and all(len(ep) == actions for ep in eps)
    for e in range(E):
        env = wb.WaveEnv(dim, design_space=ds, source=wb.Source(shapes[e], 1000.0), rng=np.random.default_rng(10 + e), **kw)
        prng = np.random.default_rng(100 + e)
        ep = wb.generate_episode(lambda env_: env_.action_space().rand(prng), env, reset=False)
        assert len(ep) == actions
        for k in range(actions):
            assert np.array_equal(ep.y[k], eps[e].y[k]) and np.array_equal(ep.t[k], eps[e].t[k])
            assert np.array_equal(ep.s[k][1], eps[e].s[k][1])                       # observation image
            assert np.array_equal(ep.a[k].table(), eps[e].a[k].table())
        assert np.array_equal(env.wave, benv.wave[e])
        s, a, t, y = wb.prepare_data(ep, 2)
        assert len(y) == actions - 1 and y[0].shape == (2 * steps + 1, 3) and t[0].shape == (2 * steps + 1,)


def test_randomized_fused_vs_exact_sweep():
    """A slice of scripts/fuzz_fused_vs_exact.py (310 random cases were green when this was written): random grid sizes
    (odd, not multiples of 4, narrower than one window), PML widths, moving designs with up to 40 cylinders, sources,
    batches -- the fused path against the exact per-stage kernels, which are bit-exact to the oracle."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_fused_vs_exact.py"), "16", "123"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    assert "FAIL" not in r.stdout and "worst" in r.stdout, r.stdout[-1500:]
    assert float(r.stdout.strip().splitlines()[-1].split()[1]) < 2e-5


@pytest.mark.parametrize("n_env", [1, 40])
def test_amortised_launches_equal_direct_launches(golden_dir, n_env):
    """waves_integrate amortises launches: one environment (a small batch) takes all steps between two saved frames in ONE
    cooperative launch with a grid barrier between steps; 40 environments replay a captured CUDA graph on calls of the same
    shape.  Consecutive env(action)-like calls (moving design, shifting tspan, frames into a device buffer) must equal the run
    that launches every kernel of every step directly, bit for bit -- including the first call after a state write (full
    interior variant) and a change of shape (re-capture)."""
    import torch
    g, dim, dyn = load_small(golden_dir)
    outs = []
    u0 = np.stack([g["u0"] * F32(1.0 + 0.01 * e) for e in range(n_env)])
    for amortise in (2, 1, 0):
        eng = make_engine(dim, dyn, n_env=n_env, dO=g["dOmega"])
        eng.set_graph(amortise)
        eng.set_state(u0)
        eng.set_source(g["shape"], float(g["freq"]))
        frames = torch.zeros((n_env, 2, 12, len(dim.y), len(dim.x)), dtype=torch.float32, device="cuda:0")
        rec = []
        for k, steps in enumerate((12, 12, 12, 8, 12)):
            t0 = F32(12 * k) * F32(1e-5)
            ts = wb.build_tspan(t0, 1e-5, steps)
            eng.set_design(g["cyl0"] if k % 2 == 0 else g["cyl1"], g["cyl1"] if k % 2 == 0 else g["cyl0"], ts[0], ts[-1])
            en, _ = eng.integrate(ts, wb.MODE_FUSED, energy=True, save_steps=[steps - 4, steps], frames=frames)
            rec.append((en.copy(), frames.cpu().numpy().copy(), eng.get_state().copy()))
            if k == 2:
                eng.set_state(rec[0][2])   # a state write invalidates the constant-field shortcut (and the graph key)
        outs.append((rec, eng.launch_count(), eng.graph_replays()))
        eng.close()
    for mode in (0, 1):
        for (ea, fa, ua), (eb, fb, ub) in zip(outs[mode][0], outs[2][0]):
            assert np.array_equal(ea, eb) and np.array_equal(fa, fb) and np.array_equal(ua, ub)
    if n_env == 1:
        assert outs[0][1] < outs[1][1] / 2 and outs[0][2] == 0, "mode 2: a small batch takes several steps per launch"
        assert outs[1][1] <= outs[2][1] and outs[1][2] == 0, "mode 1: one launch per step, no graph"
    else:
        assert outs[0][1] == outs[1][1] == outs[2][1] and outs[0][2] == outs[1][2] == 5, "a replay must account for the launches it stands for"


def test_decimated_trajectory_capture(golden_dir):
    """U-only trajectories for renderers (src/plot.jl:25) with waves_set_traj_stride: every k-th frame, batch of 2 envs."""
    g, dim, dyn = load_small(golden_dir)
    ny, nx = len(dim.y), len(dim.x)
    eng = make_engine(dim, dyn, n_env=2, dO=g["dOmega"])
    eng.set_source(g["shape"], float(g["freq"]))
    eng.set_design(g["cyl0"], g["cyl1"], g["tspan"][0], g["tspan"][-1], env=0)
    ts = wb.build_tspan(0.0, 1e-5, 20)
    u0 = np.stack([g["u0"], 0.5 * g["u0"]])
    eng.set_state(u0)
    full_t, full_i = np.empty((2, 21, ny, nx), F32), np.empty((2, 21, ny, nx), F32)
    eng.integrate(ts, wb.MODE_FUSED, energy=False, u_tot=full_t, u_inc=full_i)
    eng.set_state(u0)
    eng.set_traj_stride(5)
    dec_t, dec_i = np.empty((2, 5, ny, nx), F32), np.empty((2, 5, ny, nx), F32)
    eng.integrate(ts, wb.MODE_FUSED, energy=False, u_tot=dec_t, u_inc=dec_i)
    assert np.array_equal(dec_t, full_t[:, ::5]) and np.array_equal(dec_i, full_i[:, ::5])
    assert np.array_equal(full_t[:, 0], u0[:, 0]) and np.array_equal(full_i[:, 0], u0[:, 6])
    # device destinations take the SM copy kernel (k_copy_blocks) instead of pitched memcpys: same bytes, saved frames included
    import torch
    eng.set_state(u0)
    dev_t, dev_i = (torch.full((2, 5, ny, nx), -1.0, dtype=torch.float32, device="cuda:0") for _ in range(2))
    dev_f = torch.full((2, 3, 12, ny, nx), -1.0, dtype=torch.float32, device="cuda:0")
    eng.integrate(ts, wb.MODE_FUSED, energy=False, save_steps=[0, 7, 20], frames=dev_f, u_tot=dev_t, u_inc=dev_i)
    state_dev = eng.get_state()
    eng.set_state(u0)
    _, host_f = eng.integrate(ts, wb.MODE_FUSED, energy=False, save_steps=[0, 7, 20])
    assert np.array_equal(dev_t.cpu().numpy(), dec_t) and np.array_equal(dev_i.cpu().numpy(), dec_i)
    assert np.array_equal(dev_f.cpu().numpy(), host_f)
    assert np.array_equal(host_f[:, 0], u0) and np.array_equal(host_f[:, 2], state_dev)
    assert np.array_equal(host_f[:, 1, 0], full_t[:, 7]) and np.array_equal(host_f[:, 1, 6], full_i[:, 7])
    eng.close()


def test_config2_twenty_actions_energy_trace():
    """BASELINE configs[1] as written (scripts/data.jl:48-55): WaveEnv with the triple-ring design space and
    RandomPosGaussianSource(x = -10, y ~ U[-10, 10]), 20 actions x 100 RK4 steps; the concatenated 2001 x 3 energy trace
    (flatten_repeated_last_dim, src/utils.jl:20-31) and the final fields against the C restatement chained the same way.
    Tolerance: the north star's 1e-4 relative L2 after all 2000 steps."""
    n, steps, actions = 700, 100, 20
    dimg, dimo = wb.TwoDim(15.0, n), wo.TwoDim.make(15.0, n)
    dyn = wo.AcousticDynamics.make(dimo, wo.WATER, 2.0, 20000.0)
    dO = F32(wo.get_dx(dimo) * wo.get_dy(dimo))
    rng = np.random.default_rng(0)
    ds = wb.build_triple_ring_design_space()
    src = wb.RandomPosGaussianSource(dimg, [-10.0, -10.0], [-10.0, 10.0], [0.3], [1.0], 1000.0, rng=rng)
    env = wb.WaveEnv(dimg, design_space=ds, source=src, integration_steps=steps, actions=actions, rng=rng)
    as_cyl = lambda d: wo.Cylinders(d.table()[:, :2], d.table()[:, 2], d.table()[:, 3])
    u = np.zeros((12, n, n), F32)
    sig_g, sig_o = [], []
    while not env.is_terminated():
        d0 = env.design
        ts, _, _, _ = env(env.action_space().rand(rng))
        u, en, _ = co.integrate(dyn, u, ts, 1e-5, dO, as_cyl(d0), as_cyl(env.design), ts[0], ts[-1], shape=src.shape, freq=1000.0)
        sig_g.append(env.signal)
        sig_o.append(en)
    sg, so = wo.flatten_repeated_last_dim(np.stack(sig_g)), wo.flatten_repeated_last_dim(np.stack(sig_o))
    assert sg.shape == (actions * steps + 1, 3) and env.time_step == actions * steps
    errs = [rel(sg[:, k], so[:, k]) for k in range(3)]
    ferr = rel(env.wave[-1], u)
    per_action = env.iter.engine.launch_count() / actions
    print(f"config 2: energy trace rel-L2 tot/inc/sc = {errs}, final fields {ferr:.2e}, {per_action:.1f} launches per env(action)")
    assert so[-1, 2] > 0 and max(errs) < TOL and ferr < TOL
    assert per_action < 112, "a single environment takes one launch per step and one energy reduction per action"


def test_config3_full_batch_1024_envs_properties():
    """BASELINE configs[2] at its full size on ONE GPU: 1024 independent 700^2 environments (52 GB of state).  The oracle cannot
    run this, so size-independent properties carry the check: (i) environments with identical inputs, 12 apart in the batch and
    therefore on different warps / work items, end bitwise identical; (ii) the first 12 environments equal a 12-environment handle
    (a different work plan): fields bit for bit, energy trace to 1e-6; (iii) an environment without design has E_sc == 0 exactly and
    E_tot == E_inc; (iv) the energies are finite and the source has put energy in."""
    import torch
    if torch.cuda.mem_get_info(0)[0] < 70e9:
        pytest.skip("needs 70 GB of free device memory")
    n, E, period, steps = 700, 1024, 12, 12
    dim = wb.TwoDim(15.0, n)
    ds = wb.build_triple_ring_design_space()
    rng = np.random.default_rng(7)
    designs = []
    for k in range(4):
        d0 = ds.rand(rng)
        d1 = ds(d0, wb.build_action_space(d0, 0.25).rand(rng))
        designs.append(None if k == 3 else (d0.table(), d1.table()))
    shapes = [wb.build_normal(dim, [[-10.0, y]], [0.3], [1.0]) for y in (-6.0, 0.5, 7.0)]
    ts = wb.build_tspan(0.0, 1e-5, steps)
    # initial states: smooth random fields, incident == total (so that an environment without design keeps them equal); the wave
    # of a source at x = -10 would need ~400 steps to reach the rings, a field that is everywhere meets them at once
    u0s = []
    for k in range(3):
        a = rng.standard_normal((6, n // 20, n // 20)).astype(F32)
        f = np.kron(a, np.ones((20, 20), F32))
        for _ in range(3):   # a few smoothing passes
            f = (f + np.roll(f, 3, 1) + np.roll(f, -3, 1) + np.roll(f, 3, 2) + np.roll(f, -3, 2)) * F32(0.2)
        u0s.append(np.ascontiguousarray(np.concatenate([f, f]) * F32(1e-3))[None])

    def run(n_env):
        eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=n_env, device=0)
        for e in range(n_env):
            eng.set_source(shapes[e % 3], 1000.0, env=e)
            eng.set_state(u0s[e % 3], env=e)
            d = designs[e % 4]
            if d is None:
                eng.set_design(None, None, 0, 0, env=e)
            else:
                eng.set_design(d[0], d[1], ts[0], ts[-1], env=e)
        for _ in range(2):   # two env(action)-like calls: the first runs the full interior variant, the second the lean one
            en, _ = eng.integrate(ts, wb.MODE_FUSED)
        picks = {e: eng.get_state(e).copy() for e in (0, 3, 5, 11, 12, 15, 515, 1020, 1023) if e < n_env}
        eng.close()
        return en, picks

    en, st = run(E)
    assert np.isfinite(en).all() and en[:, -1, 1].min() > 0
    for e in range(period, E):
        assert np.array_equal(en[e], en[e % period]), f"env {e} differs from env {e % period}"
    assert np.array_equal(st[12], st[0]) and np.array_equal(st[15], st[3]) and np.array_equal(st[515], st[11])
    assert np.array_equal(st[1020], st[0]) and np.array_equal(st[1023], st[3])
    assert np.array_equal(st[3][:6], st[3][6:]), "no design: total field == incident field"
    assert np.all(en[3, :, 2] == 0.0) and np.array_equal(en[3, :, 0], en[3, :, 1])
    assert en[0, -1, 2] > 0 and not np.array_equal(st[0][0], st[0][6]), "a design scatters"
    en12, st12 = run(period)
    # (fields are independent of the work plan; the energy sums are taken per work item, so their last bits follow the plan)
    assert np.abs(en12 - en[:period]).max() <= 1e-6 * np.abs(en12).max()
    for e in (0, 3, 5, 11):
        assert np.array_equal(st12[e], st[e])
