"""CPU tests of the 1-D latent dynamics: the oracle against independent statements of the same
equations, and the product's kernel source (csrc/latent_core.cuh) executed by the host emulation of the CUDA execution
model (tests/emu/) against the oracle.  The emulation is test infrastructure: it exists because the build container has
no GPU; the GPU parity tests proper are in tests/test_gpu_zlatent.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import latent_adjoint_oracle as lao
from oracle import latent_oracle as lo
from oracle import waves_oracle as wo
from latent_cases import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32
fp = C.POINTER(C.c_float)


# ---------------------------------------------------------------------------------------------------------------
# oracle
# ---------------------------------------------------------------------------------------------------------------
def test_pml_1d_matches_reference_constructor_properties():
    """build_pml(::OneDim) (src/pml.jl:6-15): zero inside, cubic ramp, `scale` at both ends (dyn.pml[[1]] = scale)."""
    dim = wo.OneDim.make(100.0, 1024)
    pml = lo.build_pml_1d(dim.x, 10.0, 10000.0)
    assert pml[0] == F32(10000.0) and pml[-1] == F32(10000.0)
    inside = np.abs(dim.x) <= 90.0
    assert np.all(pml[inside] == 0) and np.all(np.diff(pml[: 60]) <= 0)
    i = 20
    r = (abs(float(dim.x[i])) - 90.0) / 10.0
    assert abs(float(pml[i]) - 1e4 * r ** 3) <= 2e-3 * 1e4 * r ** 3


def test_linear_interp_restatement():
    """linear_interp (src/utils.jl:70-86): piecewise linear inside the knots, the last knot included, 0 outside."""
    rng = np.random.default_rng(1)
    X = np.array([[0.0, 1.0, 2.0, 4.0], [1.0, 1.5, 2.0, 2.5]], F32)
    Y = rng.standard_normal((2, 4, 5)).astype(F32)
    for q in ([0.25, 1.75], [1.0, 2.5], [4.0, 1.0], [3.0, 2.25]):
        q = np.array(q, F32)
        got = lo.linear_interp(X, Y, q)
        for b in range(2):
            want = np.array([np.interp(float(q[b]), X[b].astype(float), Y[b, :, i].astype(float)) for i in range(5)])
            np.testing.assert_allclose(got[b], want, rtol=2e-6, atol=2e-6)
    out = lo.linear_interp(X, Y, np.array([5.0, 0.5], F32))   # outside every segment: masks are all false
    assert np.all(out == 0)


def test_rhs_matches_independent_statement_of_test_pinn():
    """test/pinn.jl:20-36 (SimpleWave) states the same 1-D right-hand side with a dense-in-form ∇: u_t = c0 c ∇v − σu (× bc),
    v_t = c0 c ∇(u + f) − σv.  Checked in float64 against the sparse matrix product."""
    cs = make_case(n=64, batch=2, steps=8, nseq=3)
    dyn, th = cs["dyn"], cs["theta"]
    w = cs["z0"]
    t = np.ascontiguousarray(cs["tspan"][:, 3])
    got = dyn(w, t, th).astype(np.float64)
    D = dyn.grad.to_scipy().toarray().astype(np.float64)
    c = lo.linear_interp(th.X, th.Y, t).astype(np.float64)
    f = th.shape.astype(np.float64) * np.sin(2 * np.pi * t.astype(np.float64) * 1000.0)[:, None]
    sig = float(dyn.pml[0]) * th.pml.astype(np.float64)
    c0 = float(dyn.c0)
    W = w.astype(np.float64)
    want = np.stack([(c0 * c * (W[:, 1] @ D.T) - sig * W[:, 0]) * dyn.bc, c0 * c * ((W[:, 0] + f) @ D.T) - sig * W[:, 1],
                     (c0 * (W[:, 3] @ D.T) - sig * W[:, 2]) * dyn.bc, c0 * ((W[:, 2] + f) @ D.T) - sig * W[:, 3]], 1)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 3e-6 * scale


def test_incident_equals_total_when_c_is_one():
    """With C ≡ 1 the total and the incident field obey the same equation (src/dynamics.jl:210-214)."""
    cs = make_case(n=64, batch=2, steps=12, nseq=3)
    th = cs["theta"]
    th.Y[:] = 1.0
    z0 = cs["z0"].copy()
    z0[:, 2:] = z0[:, :2]
    z = lo.integrate(cs["dyn"], z0, cs["tspan"], th, cs["dt"])
    np.testing.assert_allclose(z[:, :, 0], z[:, :, 2], rtol=0, atol=2e-6 * np.abs(z).max())
    e = lo.compute_latent_energy(z, wo.get_dx(cs["dim"]))
    assert e.shape == (2, 3, 13) and np.all(e[:, 2] <= 1e-9 * e[:, 0].max())


# ---------------------------------------------------------------------------------------------------------------
# kernel source under the host emulation of the CUDA execution model
# ---------------------------------------------------------------------------------------------------------------
class LatentP(C.Structure):
    """struct LatentP (csrc/latent_core.cuh)."""
    _fields_ = [("n", C.c_int), ("batch", C.c_int), ("steps", C.c_int), ("nseq", C.c_int),
                ("c0", C.c_float), ("dt", C.c_float), ("hdt", C.c_float), ("pml_scale", C.c_float), ("freq", C.c_float),
                ("dx", C.c_float), ("gf", C.c_float * 3), ("gc", C.c_float * 2), ("gl", C.c_float * 3),
                ("z0", fp), ("tspan", fp), ("X", fp), ("Y", fp), ("shape", fp), ("pml", fp), ("z", fp), ("energy", fp),
                ("z_last", fp), ("compat", C.c_int), ("zt", fp), ("w_energy", fp), ("dL_dz", fp), ("g_z0", fp), ("g_Y", fp),
                ("g_shape", fp), ("g_pml", fp)]


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "latent_emu.cpp")
    so = os.path.join(ROOT, "tests", "emu", "liblatent_emu.so")
    core = os.path.join(ROOT, "waves.jl_b200", "csrc", "latent_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                               "-fno-fast-math", "-o", so, src])
    L = C.CDLL(so)
    assert L.emu_sizeof_latentp() == C.sizeof(LatentP)
    return L


def _params(cs, **ptrs):
    dyn, th = cs["dyn"], cs["theta"]
    g = dyn.grad
    p = LatentP(n=cs["n"], batch=cs["batch"], steps=cs["steps"], nseq=cs["nseq"], c0=float(dyn.c0), dt=float(cs["dt"]),
                hdt=float(F32(0.5) * cs["dt"]), pml_scale=float(dyn.pml[0]), freq=float(th.freq),
                dx=float(wo.get_dx(cs["dim"])))
    p.gf[:], p.gc[:], p.gl[:] = list(g.first), list(g.central), list(g.last)
    keep = []
    for name, arr in dict(z0=cs["z0"], tspan=cs["tspan"], X=th.X, Y=th.Y, shape=th.shape, pml=th.pml, **ptrs).items():
        if arr is None:
            continue
        assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"]
        keep.append(arr)
        setattr(p, name, arr.ctypes.data_as(fp))
    return p, keep


@pytest.mark.parametrize("kernel,n,nt,knots,steps", [
    ("generic", 96, 96, "actions", 12), ("generic", 200, 64, "actions", 12), ("generic", 70, 96, "partial", 12),
    ("r1", 96, 96, "actions", 12), ("r1", 70, 96, "partial", 12), ("r1", 100, 128, "repeated", 12), ("r1", 170, 192, "actions", 12),
    ("r1", 33, 64, "actions", 520),
    ("r2", 96, 64, "actions", 12), ("r2", 70, 64, "partial", 12), ("r2", 100, 64, "repeated", 12), ("r2", 172, 96, "actions", 12),
    ("r2", 34, 32, "actions", 520), ("r2", 4, 32, "actions", 12), ("r2", 320, 160, "actions", 12)])
def test_emulated_forward_kernel_is_bit_exact(emu, kernel, n, nt, knots, steps):
    """k_latent_integrate / k_latent_integrate_r1 == oracle integrate, bit for bit (fields), energies to 1e-6; threads ==
    elements, a strided ownership (n > threads, generic kernel only), idle threads (threads > n), knots that leave stage
    times outside every segment, repeated knots (two masks true at once: no cursor), and more steps than one table of
    source factors holds."""
    cs = make_case(n=n, batch=2, steps=steps, nseq=5 if steps > 100 else 4, seed=n, knots=knots)
    z = np.full((cs["steps"] + 1, cs["batch"], 4, n), np.nan, F32)
    e = np.full((cs["batch"], 3, cs["steps"] + 1), np.nan, F32)
    last = np.full((cs["batch"], 4, n), np.nan, F32)
    p, keep = _params(cs, z=z, energy=e, z_last=last)
    {"generic": emu.emu_latent_integrate, "r1": emu.emu_latent_integrate_r1, "r2": emu.emu_latent_integrate_r2}[kernel](
        C.byref(p), nt)
    want = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    assert np.isfinite(want).all() and np.abs(want[-1] - want[0]).max() > 1e-3      # the case does something
    assert np.array_equal(z, want)
    assert np.array_equal(last, want[-1])
    we = lo.compute_latent_energy(want, wo.get_dx(cs["dim"]))
    np.testing.assert_allclose(e, we, rtol=1e-6, atol=1e-6 * we.max())


@pytest.mark.parametrize("kernel", ["generic", "r1", "r2"])
def test_emulated_forward_kernel_energy_only(emu, kernel):
    """With z == NULL only the energies and the last state leave the kernel."""
    cs = make_case(n=64, batch=2, steps=8, nseq=3, seed=5)
    e = np.full((2, 3, 9), np.nan, F32)
    last = np.full((2, 4, 64), np.nan, F32)
    p, keep = _params(cs, energy=e, z_last=last)
    {"generic": emu.emu_latent_integrate, "r1": emu.emu_latent_integrate_r1, "r2": emu.emu_latent_integrate_r2}[kernel](
        C.byref(p), 64)
    want = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    assert np.array_equal(last, want[-1])
    np.testing.assert_allclose(e, lo.compute_latent_energy(want, wo.get_dx(cs["dim"])), rtol=1e-6)


@pytest.mark.parametrize("compat", [0, 1])
@pytest.mark.parametrize("kernel,n,nt,knots", [("generic", 48, 64, "actions"), ("generic", 80, 32, "actions"),
                                               ("r1", 48, 64, "actions"), ("r1", 65, 96, "actions"), ("r1", 130, 160, "actions"),
                                               ("r1", 70, 96, "partial"), ("r1", 40, 64, "repeated"),
                                               ("generic", 40, 64, "repeated"),
                                               ("r2", 48, 32, "actions"), ("r2", 66, 64, "actions"), ("r2", 132, 96, "actions"),
                                               ("r2", 70, 64, "partial"), ("r2", 40, 32, "repeated"), ("r2", 4, 32, "actions"),
                                               ("r2", 8, 32, "actions"), ("r2", 260, 160, "actions"), ("r2", 130, 96, "actions")])
def test_emulated_adjoint_kernel_matches_autograd(emu, compat, kernel, n, nt, knots):
    """k_latent_adjoint / k_latent_adjoint_r1 against torch float64 reverse-mode over the unrolled oracle trajectory (exact
    mode) and a literal transcription of the reference loop (compat mode, src/dynamics.jl:101-115): z0, C.Y, F.shape and
    PML gradients; the register kernel also against the generic one."""
    steps = 8
    cs = make_case(n=n, batch=2, steps=steps, nseq=4 if knots == "repeated" else 3, seed=11 + n, knots=knots)
    rng = np.random.default_rng(3)
    T = cs["steps"] + 1
    z = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    w_energy = rng.standard_normal((2, 3, T)).astype(F32)
    dL_dz = (1e-2 * rng.standard_normal(z.shape)).astype(F32)

    def run(fn):
        g = dict(z0=np.full((2, 4, n), np.nan, F32), Y=np.zeros((2, cs["nseq"], n), F32), shape=np.full((2, n), np.nan, F32),
                 pml=np.full((2, n), np.nan, F32))
        p, keep = _params(cs, zt=z, w_energy=w_energy, dL_dz=dL_dz, g_z0=g["z0"], g_Y=g["Y"], g_shape=g["shape"],
                          g_pml=g["pml"])
        p.compat = compat
        fn(C.byref(p), nt)
        return g

    got = run({"generic": emu.emu_latent_adjoint, "r1": emu.emu_latent_adjoint_r1, "r2": emu.emu_latent_adjoint_r2}[kernel])
    want = lao.adjoint_truth(cs, w_energy, dL_dz, compat=bool(compat))
    for name in ("z0", "Y", "shape", "pml"):
        err = np.linalg.norm(got[name] - want[name]) / np.linalg.norm(want[name])
        assert err < 1e-4, (name, err)
    if kernel != "generic":
        gen = run(emu.emu_latent_adjoint)
        for name in ("z0", "Y", "shape", "pml"):
            err = np.linalg.norm(got[name] - gen[name]) / np.linalg.norm(gen[name])
            assert err < 2e-5, (name, err)


@pytest.mark.parametrize("compat", [0, 1])
def test_emulated_adjoint_r1_many_steps_and_segments(emu, compat):
    """More steps than one table of source factors holds (chunk switches in reverse order), several segments of C (the
    register accumulators of dL/dY are flushed at every segment change), energy cotangent only."""
    n, steps = 34, 520
    cs = make_case(n=n, batch=1, steps=steps, nseq=5, seed=4)
    rng = np.random.default_rng(5)
    z = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], cs["theta"], cs["dt"])
    w_energy = rng.standard_normal((1, 3, steps + 1)).astype(F32)

    def run(fn):
        g = dict(z0=np.full((1, 4, n), np.nan, F32), Y=np.zeros((1, 5, n), F32), shape=np.full((1, n), np.nan, F32),
                 pml=np.full((1, n), np.nan, F32))
        p, keep = _params(cs, zt=z, w_energy=w_energy, g_z0=g["z0"], g_Y=g["Y"], g_shape=g["shape"], g_pml=g["pml"])
        p.compat = compat
        fn(C.byref(p), 64)
        return g

    gen = run(emu.emu_latent_adjoint)
    for fn in (emu.emu_latent_adjoint_r1, emu.emu_latent_adjoint_r2):
        got = run(fn)
        for name in ("z0", "Y", "shape", "pml"):
            err = np.linalg.norm(got[name] - gen[name]) / np.linalg.norm(gen[name])
            assert err < 5e-5, (name, err)
            assert np.linalg.norm(gen[name]) > 0


def test_kernels_are_race_free_under_thread_sanitizer():
    """The six latent kernels under the host emulation, built with -fsanitize=thread: ThreadSanitizer orders accesses by
    the barriers (vector clocks, independent of timing), so a missing __syncthreads() between a shared-memory write and a
    neighbour's read is reported even if the interleaving that breaks it never happens in this run."""
    src = os.path.join(ROOT, "tests", "emu", "latent_tsan_main.cpp")
    exe = os.path.join(ROOT, "tests", "emu", "latent_tsan")
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", os.devnull], input=b"int main(){return 0;}")
    if probe.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-ffp-contract=off", "-o", exe, src])
    for args in (["70", "96", "6", "1"], ["65", "96", "5", "0"], ["200", "64", "4", "1"], ["130", "160", "5", "0"],
                 ["33", "64", "300", "1"], ["34", "64", "300", "1"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr + r.stdout, (args, (r.stderr + r.stdout)[-2000:])


def test_latent_oracle_matches_reference_vectors():
    """Pins oracle/latent_oracle.py to the Julia reference; skips while tests/golden/ref_latent_final_and_energy.f32 is absent
    (tests/golden/make_reference_vectors.jl, part 3, has never been run: no Julia here)."""
    f = os.path.join(ROOT, "tests", "golden", "ref_latent_final_and_energy.f32")
    if not os.path.exists(f):
        pytest.skip("no vectors from the Julia reference (run tests/golden/make_reference_vectors.jl in a Waves.jl checkout)")
    n, B, steps = 256, 2, 40
    dim = wo.OneDim.make(100.0, n)
    dyn = lo.LatentDynamics.make(dim, wo.WATER, 10.0, 10000.0)
    x = dim.x
    tspan = np.stack([wo.build_tspan(F32(0.0), 1e-5, steps), wo.build_tspan(F32(1e-3), 1e-5, steps)])
    X = np.ascontiguousarray(tspan[:, [0, 20, 40]])
    Y = np.stack([np.stack([(F32(1.0) + F32(0.1) * F32(k) + F32(0.2) * np.sin(F32(0.05) * F32(b) * x)).astype(F32)
                            for k in (1, 2, 3)]) for b in (1, 2)])
    shape = np.stack([np.exp(-(x + F32(20)) ** 2 / F32(100)), np.exp(-(x - F32(10)) ** 2 / F32(150))]).astype(F32)
    pml = np.stack([dyn.pml / dyn.pml.max()] * 2).astype(F32)
    z0 = np.zeros((B, 4, n), F32)
    z0[0, 0], z0[0, 2] = np.exp(-x ** 2 / F32(200)), np.exp(-(x - F32(5)) ** 2 / F32(300))
    z0[1, 0] = np.exp(-(x + F32(30)) ** 2 / F32(250))
    z0[1, 2] = z0[1, 0]
    z = lo.integrate(dyn, z0, tspan, lo.LatentTheta(X=X, Y=Y, shape=shape, freq=F32(1000.0), pml=pml), F32(1e-5))
    e = lo.compute_latent_energy(z, wo.get_dx(dim))
    ref = np.fromfile(f, F32)
    ref_z, ref_e = ref[: B * 4 * n].reshape(B, 4, n), ref[B * 4 * n:].reshape(B, 3, steps + 1)
    rel = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64))  # noqa: E731
    assert rel(z[-1], ref_z) < 1e-4 and rel(e, ref_e) < 1e-4


def test_randomized_agreement_of_all_kernels(emu):
    """A slice of scripts/fuzz_latent_emu.py: random sizes, thread counts, knot layouts, modes -- the three forward kernels
    equal the oracle bit for bit, the register reverse kernels equal the generic one to float32 rounding."""
    r = subprocess.run([os.sys.executable, os.path.join(ROOT, "scripts", "fuzz_latent_emu.py"), "7", "16"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
