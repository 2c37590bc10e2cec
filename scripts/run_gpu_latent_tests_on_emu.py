"""Runs the GPU test functions of tests/test_gpu_zlatent.py against the host emulation (mock LatentIntegrator) to check the
TEST code itself (shapes, flags, tolerances) before it first runs on a B200."""
import ctypes as C, os, sys, inspect, itertools
import numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import waves_b200 as wb
import test_latent_cpu as T
from oracle import waves_oracle as wo
L = C.CDLL(ROOT+'/tests/emu/liblatent_emu.so')
F32=np.float32

class MockIntegrator:
    def __init__(self, dyn, dt, device=0):
        self.dyn, self.dt, self.n = dyn, F32(dt), len(dyn.dim.x)
        if 21*self.n*4 > 227*1024: raise wb.WavesError("n needs too much shared memory")
        self.variant = 0; self.launches = 0
    def set_variant(self, v): self.variant = int(v)
    def set_generic(self, on): self.variant = 1 if on else 0
    def launch_count(self): return self.launches
    def _cs(self, z0, tspan, theta):
        import oracle.latent_oracle as lo
        Cint, Fsrc, pml = theta
        dim = wo.OneDim(np.ascontiguousarray(self.dyn.dim.x, F32))
        dyn = lo.LatentDynamics.make(dim, self.dyn.c0, self.dyn.pml_width, self.dyn.pml_scale)
        th = lo.LatentTheta(X=np.ascontiguousarray(Cint.X,F32), Y=np.ascontiguousarray(Cint.Y,F32), shape=None if Fsrc.shape is None else np.ascontiguousarray(Fsrc.shape,F32), freq=F32(Fsrc.freq), pml=np.ascontiguousarray(pml,F32))
        batch = th.Y.shape[0]
        return dict(dim=dim, dyn=dyn, z0=z0, tspan=np.ascontiguousarray(tspan,F32), theta=th, dt=self.dt, steps=tspan.shape[1]-1, n=self.n, batch=batch, nseq=th.X.shape[1])
    def __call__(self, z0, tspan, theta, want_z=True, want_energy=False, out_z=None, out_energy=None):
        cs = self._cs(np.ascontiguousarray(z0,F32), tspan, theta)
        B, S, n = cs["batch"], cs["steps"], self.n
        z = np.empty((S+1,B,4,n),F32) if want_z else None
        e = np.empty((B,3,S+1),F32) if want_energy else None
        last = None if want_z else np.empty((B,4,n),F32)
        p, keep = T._params(cs, z=z, energy=e, z_last=last)
        fast = self.variant != 1 and n <= 1024 and cs["nseq"] <= 64
        pair = fast and (self.variant & 2) and n % 2 == 0 and n >= 4
        nt = ((n+31)//32)*32
        if pair: L.emu_latent_integrate_r2(C.byref(p), ((n//2+31)//32)*32)
        elif fast: L.emu_latent_integrate_r1(C.byref(p), nt)
        else: L.emu_latent_integrate(C.byref(p), min(nt,1024))
        self.launches += 1
        out = z if want_z else last
        return (out, e) if want_energy else out
    def adjoint(self, z, tspan, theta, w_energy=None, dL_dz=None, mode=0):
        cs = self._cs(np.ascontiguousarray(z[0],F32), tspan, theta)
        B, n = cs["batch"], self.n
        if not ((self.variant & 4) and n <= 1024) and 39*n*4 > 227*1024: raise wb.WavesError("shared memory")
        g = {"z0": np.empty((B,4,n),F32), "Y": np.zeros((B,cs["nseq"],n),F32), "shape": np.empty((B,n),F32), "pml": np.empty((B,n),F32)}
        p, keep = T._params(cs, zt=np.ascontiguousarray(z,F32), w_energy=None if w_energy is None else np.ascontiguousarray(w_energy,F32), dL_dz=None if dL_dz is None else np.ascontiguousarray(dL_dz,F32), g_z0=g["z0"], g_Y=g["Y"], g_shape=g["shape"], g_pml=g["pml"])
        p.compat = int(mode)
        fast = (self.variant & 4) and self.variant != 1 and n <= 1024
        pair = fast and (self.variant & 2) and n % 2 == 0 and n >= 4
        nt = ((n+31)//32)*32
        if pair: L.emu_latent_adjoint_r2(C.byref(p), ((n//2+31)//32)*32)
        elif fast: L.emu_latent_adjoint_r1(C.byref(p), nt)
        else: L.emu_latent_adjoint(C.byref(p), min(nt,1024))
        return g

wb.LatentIntegrator = MockIntegrator
import test_gpu_zlatent as G
G.wb.LatentIntegrator = MockIntegrator
import pytest
names = sys.argv[1:] or [n for n in dir(G) if n.startswith('test_')]
for name in names:
    fn = getattr(G, name)
    marks = [m for m in getattr(fn, 'pytestmark', []) if m.name == 'parametrize']
    combos = [dict()]
    for m in reversed(marks):
        keys = [k.strip() for k in m.args[0].split(',')]
        new = []
        for c in combos:
            for vals in m.args[1]:
                vals = vals if isinstance(vals, (tuple, list)) and len(keys) > 1 else (vals,)
                d = dict(c); d.update(dict(zip(keys, vals))); new.append(d)
        combos = new
    for kw in combos:
        try:
            fn(**kw); print('PASS', name, kw, flush=True)
        except Exception as ex:
            print('FAIL', name, kw, type(ex).__name__, str(ex)[:300], flush=True)
