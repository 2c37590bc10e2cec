"""ctypes loader for oracle/libwaves_oracle.so (the C restatement; TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libwaves_oracle.so")
    src = os.path.join(_HERE, "waves_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libwaves_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.wo_rhs.restype = C.c_int
        _LIB.wo_integrate.restype = C.c_int
        _LIB.wo_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(fp)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def grad8(g) -> np.ndarray:
    """first(3) central(2) last(3) coefficient rows of waves_oracle.Gradient."""
    return np.concatenate([g.first, g.central, g.last]).astype(np.float32)


def _design_args(design0, design1):
    if design0 is None:
        return -1, None, None
    def flat(d):
        cy = d.all_cylinders() if hasattr(d, "all_cylinders") else d
        return np.ascontiguousarray(np.concatenate([cy.pos, cy.r[:, None], cy.c[:, None]], axis=1), dtype=np.float32)
    a, b = flat(design0), flat(design1 if design1 is not None else design0)
    return a.shape[0], a, b


def rhs(dyn, state, t, design0=None, design1=None, ti=0.0, tf=0.0, cplane=None, shape=None, freq=0.0):
    ny, nx = state.shape[1:]
    n, c0v, c1v = _design_args(design0, design1)
    g8 = grad8(dyn.grad)
    sig = _f32(dyn.pml)
    x, y = _f32(dyn.dim.x), _f32(dyn.dim.y)
    st = _f32(state)
    k = np.empty_like(st)
    cp, sh = _f32(cplane), _f32(shape)
    lib().wo_rhs(C.c_int(nx), C.c_int(ny), _p(g8), _p(sig), C.c_float(dyn.c0), _p(x), _p(y), C.c_int(n), _p(c0v),
                 _p(c1v), C.c_float(ti), C.c_float(tf), _p(cp), _p(sh), C.c_float(freq), C.c_float(t), _p(st), _p(k))
    return k


def integrate(dyn, state, tspan, dt, dOmega, design0=None, design1=None, ti=0.0, tf=0.0, cplane=None, shape=None,
              freq=0.0, save_steps=()):
    """Returns (final_state, energy(steps+1,3), frames(nsave,12,ny,nx))."""
    ny, nx = state.shape[1:]
    n, c0v, c1v = _design_args(design0, design1)
    g8 = grad8(dyn.grad)
    sig = _f32(dyn.pml)
    x, y = _f32(dyn.dim.x), _f32(dyn.dim.y)
    st = _f32(state).copy()
    ts = _f32(tspan)
    steps = len(ts) - 1
    energy = np.zeros((steps + 1, 3), dtype=np.float32)
    ss = np.ascontiguousarray(save_steps, dtype=np.int32)
    frames = np.zeros((len(ss), 12, ny, nx), dtype=np.float32)
    cp, sh = _f32(cplane), _f32(shape)
    lib().wo_integrate(C.c_int(nx), C.c_int(ny), _p(g8), _p(sig), C.c_float(dyn.c0), C.c_float(dt), _p(x), _p(y),
                       C.c_int(n), _p(c0v), _p(c1v), C.c_float(ti), C.c_float(tf), _p(cp), _p(sh), C.c_float(freq),
                       _p(ts), C.c_int(steps), _p(st), C.c_float(dOmega), _p(energy), ss.ctypes.data_as(ip),
                       C.c_int(len(ss)), _p(frames))
    return st, energy, frames


def num_threads() -> int:
    return int(lib().wo_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C restatement (torchrun exports OMP_NUM_THREADS=1: the reference arm of bench.py overrides it)."""
    lib().wo_set_num_threads(C.c_int(int(n)))
    return num_threads()
