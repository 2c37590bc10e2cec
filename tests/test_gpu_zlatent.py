"""GPU parity tests of the 1-D latent dynamics through the C ABI (`waves_latent_*`): the kernels of
csrc/latent_core.cuh against oracle/latent_oracle.py (forward: bit-exact fields, energies to 1e-6) and
oracle/latent_adjoint_oracle.py (reverse: relative L2 <= 1e-4 against float64 autograd / the literal reference loop)."""
import numpy as np
import pytest

import waves_b200 as wb
from latent_cases import make_case
from oracle import latent_adjoint_oracle as lao
from oracle import latent_oracle as lo
from oracle import waves_oracle as wo

pytestmark = pytest.mark.gpu
F32 = np.float32


def _integrator(cs, pml_width=10.0, pml_scale=10000.0):
    dyn = wb.LatentDynamics(wb.OneDim(cs["dim"].x), cs["dyn"].c0, pml_width, pml_scale)
    return wb.LatentIntegrator(dyn, cs["dt"])


def _theta(cs):
    th = cs["theta"]
    return [wb.LinearInterpolation(th.X, th.Y), wb.LatentSource(th.shape, th.freq), th.pml]


def test_host_constructors_match_oracle():
    cs = make_case(n=1024, batch=1, steps=4, nseq=3)
    dim = wb.OneDim(100.0, 1024)
    assert np.array_equal(dim.x, cs["dim"].x)
    assert np.array_equal(wb.build_pml_1d(dim, 10.0, 10000.0), cs["dyn"].pml)


@pytest.mark.parametrize("generic", [0, 1, 8])     # LATENT_AUTO (pair register kernel), LATENT_GENERIC, LATENT_SINGLE
@pytest.mark.parametrize("n,knots,steps", [(96, "actions", 12), (200, "partial", 12), (100, "repeated", 12), (1500, "actions", 12),
                                           (64, "actions", 600)])
def test_forward_bit_exact_small(n, knots, steps, generic):
    """Register fast path and generic kernel: threads == elements (padded to a warp), idle threads, knots that are not
    increasing (no cursor), n > 1024 (strided ownership: always the generic kernel), more steps than one factor table."""
    cs = make_case(n=n, batch=3, steps=steps, nseq=7 if steps > 100 else 4, seed=n, knots=knots)
    it = _integrator(cs)
    it.set_variant(generic)
    z, e = it(cs["z0"], cs["tspan"], _theta(cs), want_energy=True)
    want = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    assert np.array_equal(z, want)
    we = lo.compute_latent_energy(want, wo.get_dx(cs["dim"]))
    np.testing.assert_allclose(e, we, rtol=1e-6, atol=1e-6 * we.max())
    assert it.launch_count() == 1


def test_forward_full_size_batch():
    """The reference's latent configuration (scripts/main.jl:120-141): 1024 elements, latent_gs = 100, one action of 100
    steps for a batch of 8; energies without the trajectory agree with energies of the stored trajectory."""
    cs = make_case(n=1024, batch=8, steps=100, nseq=3, seed=7)
    it = _integrator(cs)
    z, e = it(cs["z0"], cs["tspan"], _theta(cs), want_energy=True)
    want = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    assert np.array_equal(z, want)
    last, e2 = it(cs["z0"], cs["tspan"], _theta(cs), want_z=False, want_energy=True)
    assert np.array_equal(last, want[-1]) and np.array_equal(e, e2)
    it.set_variant(wb.LATENT_GENERIC)
    zg, eg = it(cs["z0"], cs["tspan"], _theta(cs), want_energy=True)
    assert np.array_equal(zg, z)
    np.testing.assert_allclose(eg, e, rtol=1e-6)
    np.testing.assert_allclose(e, lo.compute_latent_energy(want, wo.get_dx(cs["dim"])), rtol=1e-6)


def test_no_source_and_c_equal_one_property():
    """Size-independent property at a large batch: with C ≡ 1 and equal initial data the total and incident fields stay
    bitwise equal up to the ulp-level difference of `c0 * 1 * x` vs `c0 * x` -- here exactly equal because c0 * 1 == c0."""
    cs = make_case(n=1024, batch=160, steps=50, nseq=2, seed=3)
    cs["theta"].Y[:] = 1.0
    z0 = cs["z0"].copy()
    z0[:, 2:] = z0[:, :2]
    it = _integrator(cs)
    th = [wb.LinearInterpolation(cs["theta"].X, cs["theta"].Y), wb.LatentSource(None, 0.0), cs["theta"].pml]
    z, e = it(z0, cs["tspan"], th, want_energy=True)
    # knots = [t0, tN]: only the last stage of the last step may leave the knot range (C = 0 there), so compare frames < N
    assert np.array_equal(z[:-1, :, 0], z[:-1, :, 2]) and np.array_equal(z[:-1, :, 1], z[:-1, :, 3])
    assert np.all(e[:, 2, :-1] == 0) and np.isfinite(z).all()


@pytest.mark.parametrize("compat", [False, True])
@pytest.mark.parametrize("n,steps,knots", [(80, 8, "actions"), (40, 8, "repeated"), (1024, 6, "actions")])
def test_adjoint_generic_kernel_matches_autograd(n, steps, knots, compat):
    """The generic reverse kernel against float64 autograd (exact mode) / the literal reference loop (compat mode)."""
    cs = make_case(n=n, batch=2, steps=steps, nseq=4 if knots == "repeated" else 3, seed=11 + n, knots=knots)
    rng = np.random.default_rng(3)
    it = _integrator(cs)
    it.set_variant(wb.LATENT_GENERIC)
    z = it(cs["z0"], cs["tspan"], _theta(cs))
    w_energy = rng.standard_normal((2, 3, steps + 1)).astype(F32)
    dL_dz = (1e-2 * rng.standard_normal(z.shape)).astype(F32)
    g = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=w_energy, dL_dz=dL_dz, mode=wb.ADJ_COMPAT if compat else wb.ADJ_EXACT)
    want = lao.adjoint_truth(cs, w_energy, dL_dz, compat=compat, z_stored=z)
    for name in ("z0", "Y", "shape", "pml"):
        err = np.linalg.norm(g[name] - want[name]) / np.linalg.norm(want[name])
        assert err < 1e-4, (name, err)


def test_errors_are_reported_not_thrown():
    cs = make_case(n=1500, batch=1, steps=4, nseq=3)
    it = _integrator(cs)
    z = it(cs["z0"], cs["tspan"], _theta(cs))
    with pytest.raises(wb.WavesError, match="shared memory"):      # n = 1500: generic reverse kernel, 39 n floats do not fit
        it.adjoint(z, cs["tspan"], _theta(cs), w_energy=np.ones((1, 3, 5), F32))
    with pytest.raises(wb.WavesError, match="shared memory"):
        wb.LatentIntegrator(wb.LatentDynamics(wb.OneDim(100.0, 4096), 1531.0, 10.0, 10000.0), 1e-5)


# ---- explicit kernel variants (round 2: all green on the B200; the pair forms are what LATENT_AUTO selects) -----------------
@pytest.mark.parametrize("n,knots,steps", [(96, "actions", 12), (70, "partial", 12), (100, "repeated", 12), (64, "actions", 600),
                                           (1024, "actions", 100)])
def test_pair_variant_forward_bit_exact(n, knots, steps):
    """LATENT_PAIR (two elements per thread, packed f32x2) == oracle, bit for bit."""
    cs = make_case(n=n, batch=3, steps=steps, nseq=7 if steps == 600 else (3 if steps == 100 else 4), seed=n, knots=knots)
    it = _integrator(cs)
    it.set_variant(wb.LATENT_PAIR)
    z, e = it(cs["z0"], cs["tspan"], _theta(cs), want_energy=True)
    want = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    assert np.array_equal(z, want)
    we = lo.compute_latent_energy(want, wo.get_dx(cs["dim"]))
    np.testing.assert_allclose(e, we, rtol=1e-6, atol=1e-6 * we.max())
    last, e2 = it(cs["z0"], cs["tspan"], _theta(cs), want_z=False, want_energy=True)
    assert np.array_equal(last, want[-1]) and np.array_equal(e, e2)


@pytest.mark.parametrize("compat", [False, True])
@pytest.mark.parametrize("n,steps,knots", [(80, 8, "actions"), (65, 8, "partial"), (40, 8, "repeated"), (1024, 6, "actions")])
def test_adjoint_fast_path_matches_autograd(n, steps, knots, compat):
    """Register reverse kernel (LATENT_ADJ_R1) and generic reverse kernel against float64 autograd / the literal reference loop, and
    against each other."""
    cs = make_case(n=n, batch=2, steps=steps, nseq=4 if knots == "repeated" else 3, seed=11 + n, knots=knots)
    rng = np.random.default_rng(3)
    it = _integrator(cs)
    z = it(cs["z0"], cs["tspan"], _theta(cs))
    w_energy = rng.standard_normal((2, 3, steps + 1)).astype(F32)
    dL_dz = (1e-2 * rng.standard_normal(z.shape)).astype(F32)
    mode = wb.ADJ_COMPAT if compat else wb.ADJ_EXACT
    it.set_variant(wb.LATENT_SINGLE | wb.LATENT_ADJ_R1)
    g = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=w_energy, dL_dz=dL_dz, mode=mode)
    it.set_variant(wb.LATENT_ADJ_R1 | wb.LATENT_PAIR)      # pair form where n is even, else the register kernel again
    gp = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=w_energy, dL_dz=dL_dz, mode=mode)
    it.set_generic(True)
    gg = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=w_energy, dL_dz=dL_dz, mode=mode)
    want = lao.adjoint_truth(cs, w_energy, dL_dz, compat=compat, z_stored=z)
    for name in ("z0", "Y", "shape", "pml"):
        for got in (g, gp, gg):
            err = np.linalg.norm(got[name] - want[name]) / np.linalg.norm(want[name])
            assert err < 1e-4, (name, err)
        for got in (g, gp):
            assert np.linalg.norm(got[name] - gg[name]) / np.linalg.norm(gg[name]) < 2e-5, name


def test_adjoint_fast_path_many_steps_energy_cotangent():
    """300 steps over three action segments (factor-table chunks in reverse order, dL/dY accumulators flushed at every
    segment change), energy cotangent only: fast path == generic kernel."""
    cs = make_case(n=1024, batch=3, steps=300, nseq=4, seed=9)
    it = _integrator(cs)
    z = it(cs["z0"], cs["tspan"], _theta(cs))
    wE = np.random.default_rng(1).standard_normal((3, 3, 301)).astype(F32)
    it.set_variant(wb.LATENT_SINGLE | wb.LATENT_ADJ_R1)
    g = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=wE)
    it.set_variant(wb.LATENT_ADJ_R1 | wb.LATENT_PAIR)
    gp = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=wE)
    it.set_generic(True)
    gg = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=wE)
    for name in ("z0", "Y", "shape", "pml"):
        assert np.linalg.norm(gg[name]) > 0
        for got in (g, gp):
            assert np.linalg.norm(got[name] - gg[name]) / np.linalg.norm(gg[name]) < 5e-5, name




def test_adjoint_accepts_float64_cotangents_and_vector_tspan():
    """converted copies of float64 / non-contiguous cotangents must outlive the call; a vector tspan is the same
    time column for every batch element (src/dynamics.jl:51-53), in the reverse pass too."""
    cs = make_case(n=128, batch=2, steps=10, nseq=3, seed=5)
    cs["tspan"] = np.ascontiguousarray(np.broadcast_to(cs["tspan"][:1], cs["tspan"].shape))
    it = _integrator(cs)
    z = it(cs["z0"], cs["tspan"][0], _theta(cs))
    assert np.array_equal(z, it(cs["z0"], cs["tspan"], _theta(cs)))
    rng = np.random.default_rng(2)
    wE = rng.standard_normal((2, 3, 11))
    gz = 1e-2 * rng.standard_normal(z.shape)
    want = it.adjoint(z, cs["tspan"], _theta(cs), w_energy=wE.astype(F32), dL_dz=gz.astype(F32))
    got = it.adjoint(z, cs["tspan"][0], _theta(cs), w_energy=wE, dL_dz=np.asfortranarray(gz))
    for name in ("z0", "Y", "shape", "pml"):
        assert np.array_equal(got[name], want[name]), name


def test_reference_adjoint_sensitivity_script():
    """scripts/adjoint_sensitivity.jl as written: OneDim(15, 1024), AcousticDynamics(dim, WATER, 5, 10000), dt = 1e-5, N = 300 steps,
    one sample, C = LinearInterpolation(t[[1, end], :], ones), F = Source(zeros, 1), PML = dyn.pml / maximum(dyn.pml), loss =
    mse(z[:, 1, 1, end], target): the gradient with respect to z0 (what `back(one(loss))` pulls through the Integrator rrule)
    against float64 autograd, with the default kernels and the generic ones."""
    n, steps = 1024, 300
    cs = make_case(n=n, batch=1, steps=steps, nseq=2, seed=3, gs=15.0, dt=1e-5, c0=wo.WATER, pml_width=5.0, pml_scale=10000.0, freq=1.0)
    th = cs["theta"]
    pml = (cs["dyn"].pml / cs["dyn"].pml.max())[None, :].astype(F32)
    cs["theta"] = lo.LatentTheta(X=np.ascontiguousarray(cs["tspan"][:, [0, -1]]), Y=np.ones_like(th.Y), shape=np.zeros_like(th.shape),
                                 freq=F32(1.0), pml=pml)
    x = cs["dim"].x.astype(np.float64)
    rng = np.random.default_rng(0)
    # z0 = emb(freq_coefs): a smooth, small superposition of sines per field (SinWaveEmbedder), 50 frequencies
    k = np.arange(1, 51)[:, None]
    z0 = (0.01 * rng.standard_normal((4, 50)) @ np.sin(np.pi * k * (x[None, :] - x[0]) / (x[-1] - x[0]))).astype(F32)
    cs["z0"] = z0[None]
    target = np.exp(-0.5 * (x / 0.3) ** 2).astype(F32)   # build_normal(x, [0], [0.3], [1])
    it = _integrator(cs, pml_width=5.0, pml_scale=10000.0)
    z = it(cs["z0"], cs["tspan"], _theta(cs))
    assert np.array_equal(z, lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"]))
    dL_dz = np.zeros_like(z)
    dL_dz[-1, 0, 0] = (2.0 / n) * (z[-1, 0, 0] - target)   # d mse(z[:, 1, 1, end], target) / dz
    want = lao.adjoint_truth(cs, None, dL_dz, compat=False, z_stored=z)
    g = it.adjoint(z, cs["tspan"], _theta(cs), dL_dz=dL_dz)
    it.set_generic(True)
    gg = it.adjoint(z, cs["tspan"], _theta(cs), dL_dz=dL_dz)
    for got in (g, gg):
        err = np.linalg.norm(got["z0"] - want["z0"]) / np.linalg.norm(want["z0"])
        assert err < 1e-4, err
    assert np.linalg.norm(want["z0"]) > 0
    it.close()
