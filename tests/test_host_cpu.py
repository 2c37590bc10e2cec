"""CPU-only tests of the product's host side: the C ABI library loads and exports every symbol the header
declares, its host-side restatements of the reference constructors agree with the oracle, the Python
mirror of the reference types behaves like the oracle's, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import waves_b200 as wb
from oracle import c_oracle as co
from oracle import waves_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "waves_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(waves_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(wb._lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/waves_b200.h but not exported"
    assert declared == set(wb._lib.SYMBOLS), "ctypes binding and header disagree"
    assert wb._lib.lib().waves_version() == 100


def test_no_oracle_in_product():
    """The product must never import or link the oracle (test infrastructure only)."""
    pkg = os.path.join(ROOT, "waves.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".jl")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f"{f} mentions the oracle"


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    dim = wb.TwoDim(15.0, 64)
    with pytest.raises(wb.WavesError, match="no CUDA device|no CPU path"):
        wb.Engine(dim.x, dim.y, 1531.0, 1e-5)


@pytest.mark.parametrize("gs,n", [(15.0, 700), (5.0, 512), (3.0, 96), (2.0, 33)])
def test_host_builders_match_oracle(gs, n):
    """Self-consistency, not an independent pin: the product's host builders (C ABI) and the oracle's are the same derivation of
    the reference's formulas typed twice (Float32 ranges evaluated in Float64, src/dims.jl:56-60; PML profile src/pml.jl:21-29;
    gradient rows src/operators.jl:10-22).  What this catches is drift between the two; what pins either is Julia
    (tests/golden/make_reference_vectors.jl).  The GPU parity tests at 700^2 run the product's OWN builders against the oracle."""
    d = wb.TwoDim(gs, n)
    o = wo.TwoDim.make(gs, n)
    assert np.array_equal(d.x, o.x) and np.array_equal(d.y, o.y)
    assert np.array_equal(wb.build_pml(d, 2.0 if gs > 4 else 0.5, 20000.0), wo.build_pml_profile(o.x, 2.0 if gs > 4 else 0.5, 20000.0))
    assert np.array_equal(wb.build_gradient(d), co.grad8(wo.build_gradient(o.x)))
    assert abs(float(wb.get_dx(d)) - float(wo.get_dx(o))) <= 2e-7 * float(wo.get_dx(o))
    mu, sg, a = np.array([[-0.4 * gs, 0.1 * gs], [0.2, -0.3]]), np.array([0.3, 0.2]), np.array([1.0, 0.5])
    assert np.array_equal(wb.build_normal(d, mu, sg, a), wo.build_normal(wo.build_grid(o), mu, sg, a))


def test_tspan_matches_oracle():
    """Self-consistency of the two statements of build_tspan (src/dynamics.jl:5-7), see test_host_builders_match_oracle."""
    for ts, steps in [(0, 100), (100, 100), (1900, 100), (37, 13)]:
        ti = F32(F32(ts) * F32(1e-5))
        assert np.array_equal(wb.build_tspan(ti, 1e-5, steps), wo.build_tspan(ti, 1e-5, steps))


def test_design_types_match_oracle():
    a, b = wb.build_triple_ring_design_space(), wo.build_triple_ring_design_space()
    assert np.array_equal(a.low.table(), np.concatenate([b.low.all_cylinders().pos, b.low.all_cylinders().r[:, None],
                                                         b.low.all_cylinders().c[:, None]], 1))
    ra, rb = np.random.default_rng(5), np.random.default_rng(5)
    da = a.rand(ra)
    # the oracle samples radii only; the mirror follows rand(::DesignSpace{Cylinders}) and draws pos, r, c
    db = wo.Cloak(wo.Cylinders(da.config.pos, da.config.r, da.config.c), wo.Cylinders(da.core.pos, da.core.r, da.core.c))
    act = wb.build_action_space(da, 0.25).rand(ra)
    acto = wo.Cylinders(act.pos, act.r, act.c)
    na, nb = a(da, act), b(db, acto)
    assert np.array_equal(na.table()[:, 2], nb.all_cylinders().r)
    ia, ib = wb.DesignInterpolator(da, na, F32(0.001), F32(0.002)), wo.DesignInterpolator(db, nb, F32(0.001), F32(0.002))
    for t in (0.0005, 0.001, 0.00137, 0.002, 0.003):
        assert np.array_equal(ia(F32(t)).table()[:, 2], ib(F32(t)).all_cylinders().r)


def test_random_pos_gaussian_source_in_bounds():
    d = wb.TwoDim(15.0, 128)
    s = wb.RandomPosGaussianSource(d, [[-10.0, -10.0]], [[-10.0, 10.0]], [0.3], [1.0], 1000.0, rng=np.random.default_rng(3))
    assert s.mu[0, 0] == -10 and -10 <= s.mu[0, 1] <= 10 and s.shape.shape == (128, 128)


def test_bench_reference_arm_prints_contract_json():
    """bench.py --impl reference: the CPU restatement timed on bounded samples; must print ONE JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-rk4-steps", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_prepare_data_mirror_matches_restatement():
    """prepare_data / flatten_repeated_last_dim (src/data.jl:35-58, src/utils.jl:20-31): host bookkeeping of the episode layer."""
    rng = np.random.default_rng(0)
    steps, actions, horizon = 10, 6, 3
    ep = wb.Episode()
    t0 = 0.0
    for k in range(actions):
        ts = wb.build_tspan(np.float32(t0), 1e-5, steps)
        t0 = ts[-1]
        y = rng.standard_normal((steps + 1, 3)).astype(np.float32)
        if ep.y:
            y[0] = ep.y[-1][-1]      # consecutive actions share their boundary frame
        ep.s.append((ts, None, None))
        ep.a.append(k)
        ep.t.append(ts)
        ep.y.append(y)
    s, a, t, y = wb.prepare_data(ep, horizon)
    to, yo = wo.prepare_data(ep.t, ep.y, horizon)
    assert len(s) == len(a) == len(t) == len(y) == actions - horizon + 1 == len(to)
    for i in range(len(t)):
        assert t[i].shape == (horizon * steps + 1,) and y[i].shape == (horizon * steps + 1, 3)
        assert np.array_equal(t[i], to[i]) and np.array_equal(y[i], yo[i])
        assert a[i] == list(range(i, i + horizon)) and s[i] is ep.s[i]
        assert np.all(np.diff(t[i]) > 0)


# ---- static check of the (never executed) Julia shim against the ctypes binding that IS exercised -------------------------
_JL_TYPES = {"Int32": C.c_int32, "UInt32": C.c_uint32, "Float32": C.c_float, "Ptr{Float32}": C.POINTER(C.c_float)}


def _julia_struct(src, name):
    body = re.search(r"struct %s\n(.*?)\nend" % name, src, re.S).group(1)
    fields = []
    for decl in re.split(r"[;\n]", body):
        decl = decl.split("#")[0].strip()
        if decl:
            fname, ftype = [x.strip() for x in decl.split("::")]
            fields.append((fname, _JL_TYPES[ftype]))
    return fields


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({":
            depth += 1
        elif ch in ")}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_shim_matches_the_ctypes_binding():
    """waves.jl_b200/julia/WavesB200.jl cannot run here (no Julia).  Its struct layouts and every ccall signature are
    compared with waves.jl_b200/_lib.py, whose structures and argument lists are exercised by the GPU tests."""
    from waves_b200 import _lib
    src = open(os.path.join(ROOT, "waves.jl_b200", "julia", "WavesB200.jl")).read()
    for jl_name, ct in (("WavesConfig", _lib.WavesConfig), ("LatentConfig", _lib.LatentConfig)):
        got = _julia_struct(src, jl_name)
        want = [(n, t) for n, t in ct._fields_]
        assert [n for n, _ in got] == [n for n, _ in want], jl_name
        for (n, tj), (_, tc) in zip(got, want):
            assert C.sizeof(tj) == C.sizeof(tc) and (tj is tc or issubclass(tj, C._Pointer) == issubclass(tc, C._Pointer)), (jl_name, n)

    kinds = {"Cint": "i32", "Cfloat": "f32", "Cstring": "str", "Ptr{Cvoid}": "ptr", "Ptr{Float32}": "ptr", "Ptr{Int32}": "ptr",
             "Ref{Ptr{Cvoid}}": "ptr", "Ref{WavesConfig}": "ptr", "Ref{LatentConfig}": "ptr", "Int64": "i64"}

    def ckind(t):
        if t in (C.c_int, C.c_int32):
            return "i32"
        if t is C.c_float:
            return "f32"
        if t is C.c_int64:
            return "i64"
        if t is C.c_char_p:
            return "str"
        return "ptr"      # c_void_p and POINTER(...)

    assert "ccall((:" not in src.split("module WavesB200")[1].replace("`ccall((:name, lib), ...)`", ""), \
        "entry points are resolved with Libdl.dlsym (a (name, lib) pair must be a constant expression)"
    calls = re.findall(r"ccall\(sym\(:(waves_[a-z0-9_]+)\),\s*(\w+),\s*\((.*?)\),?\s*\n?\s*(?:h\.ptr|h_lib|cfg|\))", src, re.S)
    seen = set()
    for name, ret, args in calls:
        res, argtypes = _lib.SYMBOLS[name]
        jl_args = [a for a in _split_top(args.replace("\n", " ")) if a]
        assert [kinds[a] for a in jl_args] == [ckind(t) for t in argtypes], (name, jl_args)
        assert kinds[ret] == ckind(res), name
        seen.add(name)
    assert {"waves_create", "waves_integrate", "waves_adjoint", "waves_set_design", "waves_observe", "waves_latent_create",
            "waves_latent_integrate", "waves_latent_adjoint"} <= seen, seen


def test_header_is_c99_and_the_library_links_from_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/waves_b200.h must compile as C99 (cgo / ccall / JNI hosts) and a plain C
    program must link against the library and get error codes -- not exceptions -- back."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "waves_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "abi_link.c"
    src.write_text(r'''
#include "waves_b200.h"
#include <stdio.h>
#include <string.h>
int main(void) {
    waves_latent_config lc; waves_latent *lh = 0; waves_config c; waves_handle *h = 0; int rc1, rc2;
    memset(&lc, 0, sizeof lc); memset(&c, 0, sizeof c);
    lc.n = 16;
    rc1 = waves_latent_create(&lc, &lh);
    printf("%d|%s\n", rc1, waves_last_error());
    rc2 = waves_create(&c, &h);
    printf("%d|%s\n", rc2, waves_last_error());
    return (waves_version() == WAVES_B200_VERSION && rc1 != 0 && rc2 != 0 && !lh && !h) ? 0 : 1;
}
''')
    exe = tmp_path / "abi_link"
    libdir = os.path.join(ROOT, "waves.jl_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-L" + libdir,
                           "-lwaves_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "waves_latent_create" in out.stdout and "waves_create" in out.stdout
