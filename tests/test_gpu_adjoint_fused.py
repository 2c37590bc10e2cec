"""Fused reverse step (kernels_adjoint_fused.cu) and checkpointed storage of waves_adjoint.

Without dL/dc -- the gradient the reference itself can produce, its cylinder mask has no derivative -- a reverse step is one
fused launch set: an RK4 step of the transposed operator on shared-memory tiles.  Checked against float64 autograd over the
unrolled trajectory / the reference loop as written (oracle/adjoint_oracle.py), against the per-stage reverse kernels, by
central finite differences of the loss computed on the GPU, and with every segment length of the checkpointed storage."""
import numpy as np
import pytest

import waves_b200 as wb
from oracle import adjoint_oracle as ao
from test_gpu_adjoint import TOL, rel, setup

pytestmark = pytest.mark.gpu
F32 = np.float32


# n = 48: general tiles only (the zero-sigma zone is narrower than a tile); n = 200 with a thin PML: interior tiles, general
# tiles and the seam between the two launches; n = 131: odd size, row pitch != nx, partial tiles
# n = 200 with a 25-cell PML: the PML ring runs on the march as well (variants 7, 8, 9), tiles only on the outer 8 cells
@pytest.mark.parametrize("n,pml_width", [(48, 0.5), (200, 0.25), (131, 0.2), (200, 0.5)])
@pytest.mark.parametrize("adj_mode", [wb.ADJ_EXACT, wb.ADJ_COMPAT])
def test_fused_reverse_matches_autograd_and_per_stage_kernels(n, pml_width, adj_mode):
    steps = 6 if adj_mode == wb.ADJ_EXACT else 5
    p, eng, ts, z0, w, aN = setup(n=n, steps=steps, pml_width=pml_width)
    L, gz, _, traj = ao.autograd_truth(p, z0, w, aN)
    want = gz if adj_mode == wb.ADJ_EXACT else ao.reference_loop_torch(p, traj, w, aN)[0]
    for fwd in (wb.MODE_EXACT, wb.MODE_FUSED):
        eng.set_state(z0[None])
        n0 = eng.launch_count()
        loss, dz0, none = eng.adjoint(ts, w, aN[None], fwd_mode=fwd, adj_mode=adj_mode, want_dc=False, ring="always")
        assert none is None and rel(dz0[0], want) < TOL, (fwd, rel(dz0[0], want))
        for f in range(12):
            assert rel(dz0[0, f], want[f]) < 5 * TOL, f"field {f}"
        assert abs(loss[0] - (L - float(np.sum(traj[-1] * aN)))) < 1e-4 * abs(L)
        assert rel(eng.get_state(0), traj[-1]) < TOL
    eng.set_state(z0[None])
    _, dz_s, _ = eng.adjoint(ts, w, aN[None], fwd_mode=wb.MODE_FUSED, adj_mode=adj_mode, want_dc=False, fused_reverse=False)
    assert rel(dz0[0], np.float64(dz_s[0])) < 1e-5
    eng.close()


@pytest.mark.parametrize("every", [1, 2, 3, 7])
@pytest.mark.parametrize("want_dc", [False, True])
def test_checkpointed_storage_gives_the_same_gradients(every, want_dc):
    """7 steps in segments of 1, 2, 3 (uneven last segment) and 7 (one segment): checkpoints + re-run segments vs everything
    stored.  The re-run starts from a restored state (full interior variant instead of the lean one), so equality is to
    rounding, not bitwise."""
    p, eng, ts, z0, w, aN = setup(n=200, steps=7, pml_width=0.25)
    outs = []
    for k in (0, every):
        eng.set_adjoint_checkpoint(k)
        eng.set_state(z0[None])
        outs.append(eng.adjoint(ts, w, aN[None], fwd_mode=wb.MODE_FUSED, adj_mode=wb.ADJ_EXACT, want_dc=want_dc) + (eng.get_state(0),))
    (l0, g0, c0, s0), (l1, g1, c1, s1) = outs
    assert rel(g1[0], np.float64(g0[0])) < 2e-6 and abs(l1[0] - l0[0]) <= 1e-6 * abs(l0[0]) and np.array_equal(s0, s1)
    if want_dc:
        assert rel(c1[0], np.float64(c0[0])) < 2e-6
    eng.close()


def _env700(n_env=1, steps=500, src=(-3.5, 2.5)):
    dim = wb.TwoDim(15.0, 700)
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=n_env)
    eng.set_source(wb.build_normal(dim, [list(src)], [0.3], [1.0]), 1000.0)   # 4 m from the ring: scattering within the run
    d0 = wb.build_triple_ring_design_space().rand(np.random.default_rng(0))
    ts = wb.build_tspan(0.0, 1e-5, steps)
    eng.set_design(d0.table(), d0.table(), ts[0], ts[-1])
    return eng, ts


def test_config5_full_size_fused_reverse_vs_per_stage_and_finite_differences():
    """BASELINE configs[4]: 700^2, triple-ring design frozen, 500 steps, L = sum_t E_sc(t).
    (i) dL/dz0 of the fused reverse sweep against the per-stage kernels in the reference's exact float32 order (forward and
    reverse), relative L2 <= 1e-4; (ii) central finite differences of L along 10 random directions, L evaluated by the fused
    forward pass on the GPU (L is quadratic in z0, so the central difference is exact up to float32 rounding of the energies)."""
    import torch
    steps = 500
    eng, ts = _env700(steps=steps)
    w = np.zeros((steps + 1, 3), F32)
    w[:, 2] = 1.0
    rng = np.random.default_rng(1)
    z0 = np.zeros((1, 12, 700, 700), F32)
    eng.set_state(z0)
    loss, gz, _ = eng.adjoint(ts, w, want_dc=False)
    eng.set_state(z0)
    _, gz_e, _ = eng.adjoint(ts, w, fwd_mode=wb.MODE_EXACT, want_dc=False, fused_reverse=False)
    err = rel(gz[0], np.float64(gz_e[0]))
    print(f"config 5: dL/dz0 fused vs exact-order per-stage kernels: rel-L2 {err:.2e}, |g| = {np.linalg.norm(np.float64(gz[0])):.3e}, loss {loss[0]:.4e}")
    assert loss[0] > 0 and np.isfinite(gz).all() and err < TOL

    def L_of(z):
        eng.set_state(z)
        en, _ = eng.integrate(ts, wb.MODE_FUSED, energy=True)
        return float(np.sum(np.float64(en[0]) * w))

    g64 = np.float64(gz[0])
    L0, u = float(loss[0]), 1e-6     # u: relative rounding of a float32 energy (per-warp float32 partial sums, float64 above)
    yy, xx = np.mgrid[0:700, 0:700]
    for k in range(10):
        # smooth random direction in every field (a few random Gaussian bumps)
        d = np.zeros((12, 700, 700), F32)
        for f in range(12):
            for _ in range(3):
                cx, cy, s = rng.uniform(80, 620), rng.uniform(80, 620), rng.uniform(8, 40)
                d[f] += rng.standard_normal() * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s)).astype(F32)
        gd = float(np.sum(g64 * d))
        # L(z0 + e d) = L0 + e <g, d> + e^2 q: a trial pair gives q, and e* = sqrt(L0 / q) balances the rounding of the two
        # terms (error of the central difference ~ u (L0 + e^2 q) / e)
        e1 = 1e-2
        q = max((L_of((z0[0] + F32(e1) * d)[None]) + L_of((z0[0] - F32(e1) * d)[None]) - 2 * L0) / (2 * e1 * e1), 1e-30)
        eps = float(np.sqrt(L0 / q))
        fd = (L_of((z0[0] + F32(eps) * d)[None]) - L_of((z0[0] - F32(eps) * d)[None])) / (2.0 * float(F32(eps)))
        bound = 1e-3 * abs(gd) + 4 * u * (L0 + eps * eps * q) / eps
        print(f"  direction {k}: <g, d> = {gd:.6e}, central difference = {fd:.6e}, eps = {eps:.2e}, |diff| = {abs(fd - gd):.2e} (bound {bound:.2e})")
        assert abs(fd - gd) < bound, (k, fd, gd, bound)
    eng.close()


def test_batched_gradients_with_checkpoints_at_full_size():
    """8 x 700^2 x 120 steps with a checkpoint every 16 steps: every environment of the batch gets the gradient of the
    single-environment run (environment 3 has its own design and source position)."""
    steps = 120
    eng, ts = _env700(n_env=8, steps=steps)
    dim = wb.TwoDim(15.0, 700)
    d3 = wb.build_triple_ring_design_space().rand(np.random.default_rng(3))
    eng.set_design(d3.table(), d3.table(), ts[0], ts[-1], env=3)
    eng.set_source(wb.build_normal(dim, [[-2.0, -3.0]], [0.3], [1.0]), 1000.0, env=3)
    w = np.zeros((steps + 1, 3), F32)
    w[:, 2] = 1.0
    w[-1, 0] = 0.5
    z0 = (np.random.default_rng(2).standard_normal((1, 12, 700, 700)) * 1e-4).astype(F32)
    import torch
    gz_d = torch.empty((8, 12, 700, 700), dtype=torch.float32, device="cuda:0")
    eng.set_adjoint_checkpoint(16)
    eng.set_state(np.repeat(z0, 8, 0))
    loss, _, _ = eng.adjoint(ts, w, want_dc=False, out_dz0=gz_d)
    g = gz_d.cpu().numpy()
    one, ts1 = _env700(steps=steps)
    one.set_state(z0)
    l1, g1, _ = one.adjoint(ts1, w, want_dc=False)
    for e in (0, 1, 2, 4, 5, 6, 7):
        assert rel(g[e], np.float64(g1[0])) < 2e-6 and abs(loss[e] - l1[0]) <= 2e-6 * abs(l1[0]), e
    assert rel(g[3], np.float64(g1[0])) > 1e-2 and np.isfinite(g).all()
    one.close()
    eng.close()


@pytest.mark.parametrize("n,pml_width", [(200, 0.25), (131, 0.2)])
@pytest.mark.parametrize("adj_mode", [wb.ADJ_EXACT, wb.ADJ_COMPAT])
def test_march_interior_equals_tile_kernels(n, pml_width, adj_mode):
    """The interior of the fused reverse step runs on the forward kernel's march (TMA ring, transposed operator, auxiliary
    cotangents accumulated in one plane) with the shared-memory tiles on the frame around it; `march=False` takes the tiles
    everywhere.  Same gradient to float32 rounding, with a moving design, energy weights on every frame and a cotangent of
    the last state in all 12 fields (the auxiliary ones included)."""
    steps = 7
    p, eng, ts, z0, w, aN = setup(n=n, steps=steps, pml_width=pml_width)
    outs = []
    for march in (True, False):
        eng.set_state(z0[None])
        l0 = eng.launch_count()
        loss, dz0, _ = eng.adjoint(ts, w, aN[None], adj_mode=adj_mode, want_dc=False, march=march)
        outs.append((loss.copy(), dz0.copy(), eng.launch_count() - l0))
    (la, ga, na), (lb, gb, nb) = outs
    assert na != nb, "the two forms must not launch the same kernels"
    assert la[0] == lb[0]
    for f in range(12):
        assert rel(ga[0, f], np.float64(gb[0, f])) < 2e-6, f"field {f}: {rel(ga[0, f], np.float64(gb[0, f]))}"
    eng.close()


@pytest.mark.parametrize("n,pml_width", [(200, 0.5), (256, 0.6)])
@pytest.mark.parametrize("adj_mode", [wb.ADJ_EXACT, wb.ADJ_COMPAT])
def test_march_pml_ring_equals_tile_kernels(n, pml_width, adj_mode):
    """With a PML of >= 16 cells the ring between the outer 8 cells of the domain and the interior rectangle runs on the march
    too (left / right strips, top / bottom strips, corners: stage_TP), the tiles keep the outer 8 cells.  Three forms of the
    same gradient: ring + interior on the march, interior only, tiles everywhere."""
    steps = 7
    p, eng, ts, z0, w, aN = setup(n=n, steps=steps, pml_width=pml_width)
    outs = []
    for kw in (dict(ring="always"), dict(ring=False), dict(march=False)):   # (a single small environment skips the ring by default)
        eng.set_state(z0[None])
        l0 = eng.launch_count()
        loss, dz0, _ = eng.adjoint(ts, w, aN[None], adj_mode=adj_mode, want_dc=False, **kw)
        outs.append((loss.copy(), dz0.copy(), eng.launch_count() - l0))
    assert outs[0][2] > outs[1][2] > outs[2][2] - 2, [o[2] for o in outs]   # 5 / 2 / 2 launches per reverse step (+ the G pass)
    assert outs[0][2] != outs[1][2], "the ring variants must have been launched"
    for k in (0, 1):
        assert outs[k][0][0] == outs[2][0][0]
        for f in range(12):
            assert rel(outs[k][1][0, f], np.float64(outs[2][1][0, f])) < 2e-6, f"form {k} field {f}: {rel(outs[k][1][0, f], np.float64(outs[2][1][0, f]))}"
    eng.close()


def test_march_pml_ring_with_cylinders_in_the_ring_and_a_batch():
    """The ring variants evaluate c^2 from the culled cylinders like the forward speed field: a moving cylinder that reaches into
    the PML ring (but not into the outer 8 cells), three environments with different designs on one handle, ring + interior on
    the march against the tiles everywhere, and every environment of the batch against its own single-environment run."""
    n, steps, dt = 256, 6, 4e-6
    dim = wb.TwoDim(2.0, n)
    ts = wb.build_tspan(2e-4, dt, steps)
    rng = np.random.default_rng(11)
    z0 = (rng.standard_normal((3, 12, n, n)) * 1e-3).astype(F32)
    w = rng.uniform(0.2, 1.0, (steps + 1, 3)).astype(F32)
    shape = wb.build_normal(dim, [[-0.8, 0.2]], [0.2], [1.0])
    designs = [(np.array([[1.30, 0.20, 0.40, 1032.0], [-0.2, -1.35, 0.30, 2100.0]], F32), np.array([[1.35, 0.15, 0.45, 1032.0], [-0.2, -1.40, 0.35, 2100.0]], F32)),
               (np.array([[0.40, 0.00, 0.45, 1032.0], [-1.3, 1.30, 0.35, 900.0]], F32), np.array([[0.45, 0.05, 0.50, 1032.0], [-1.35, 1.35, 0.30, 900.0]], F32)),
               None]

    def make(n_env, which):
        eng = wb.Engine(dim.x, dim.y, wb.WATER, dt, 0.6, 20000.0, n_env=n_env)
        for e, k in enumerate(which):
            eng.set_source(shape, 1000.0, env=e)
            if designs[k] is None:
                eng.set_design(None, None, 0, 0, env=e)
            else:
                eng.set_design(designs[k][0], designs[k][1], ts[0], ts[-1], env=e)
        return eng

    eng = make(3, (0, 1, 2))
    outs = []
    for kw in (dict(ring="always"), dict(march=False)):
        eng.set_state(z0)
        outs.append(eng.adjoint(ts, w, want_dc=False, **kw)[:2])
    for e in range(3):
        assert rel(outs[0][1][e], np.float64(outs[1][1][e])) < 2e-6 and outs[0][0][e] == outs[1][0][e], e
    eng.close()
    for e in range(3):
        one = make(1, (e,))
        one.set_state(z0[e:e + 1])
        l1, g1, _ = one.adjoint(ts, w, want_dc=False, ring="always")
        assert np.array_equal(g1[0], outs[0][1][e]) and l1[0] == outs[0][0][e], e
        one.close()
