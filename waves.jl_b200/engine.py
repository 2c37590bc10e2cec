"""Thin object wrapper over the C ABI handle (include/waves_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

F32 = np.float32
MODE_FUSED, MODE_EXACT = 0, 1
STEP_ASYNC = 0x100
ADJ_EXACT, ADJ_COMPAT, ADJ_STAGEWISE, ADJ_TILES, ADJ_NO_RING, ADJ_RING = 0, 1, 0x100, 0x200, 0x400, 0x800


def _ptr(a):
    """Raw address of a NumPy array (host) or a torch tensor (host or CUDA); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype in (np.float32, np.int32) and a.flags["C_CONTIGUOUS"], "need a contiguous float32/int32 array"
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


class Engine:
    """One waves_handle: `n_env` independent environments (or one slab of a large grid) on one GPU."""

    def __init__(self, x, y, c0, dt, pml_width=2.0, pml_scale=20000.0, n_env=1, device=0, sigma=None, grad8=None,
                 d_omega=0.0, ny_local=None, row0=0):
        self._x = np.ascontiguousarray(x, F32)
        self._y = np.ascontiguousarray(y, F32)
        self._sigma = None if sigma is None else np.ascontiguousarray(sigma, F32)
        self._grad8 = None if grad8 is None else np.ascontiguousarray(grad8, F32)
        self.nx, self.ny_global = len(self._x), len(self._y)
        self.ny = int(ny_local) if ny_local is not None else self.ny_global
        self.n_env, self.device = int(n_env), int(device)
        cfg = _lib.WavesConfig(
            nx=self.nx, ny=self.ny, n_env=self.n_env, device=self.device, c0=float(c0), dt=float(dt),
            pml_width=float(pml_width), pml_scale=float(pml_scale), x=self._x.ctypes.data_as(_lib.fp),
            y=self._y.ctypes.data_as(_lib.fp),
            sigma=None if self._sigma is None else self._sigma.ctypes.data_as(_lib.fp),
            grad8=None if self._grad8 is None else self._grad8.ctypes.data_as(_lib.fp),
            d_omega=float(d_omega), ny_global=self.ny_global, row0=int(row0), flags=0)
        h = C.c_void_p()
        check(_lib.lib().waves_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().waves_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # state -------------------------------------------------------------------
    def set_state(self, u12, env=-1):
        if isinstance(u12, np.ndarray):
            u12 = np.ascontiguousarray(u12, F32)
        check(_lib.lib().waves_set_state(self._h, env, _ptr(u12)))

    def get_state(self, env=-1, out=None):
        n = self.n_env if env < 0 else 1
        if out is None:
            out = np.empty((n, 12, self.ny, self.nx), dtype=F32)
        check(_lib.lib().waves_get_state(self._h, env, _ptr(out)))
        return out if env < 0 else (out[0] if isinstance(out, np.ndarray) and out.ndim == 4 else out)

    # θ = [C, F] --------------------------------------------------------------
    def set_source(self, shape, freq, env=-1):
        check(_lib.lib().waves_set_source(self._h, env, _ptr(shape), C.c_float(freq)))

    def set_design(self, cyl0, cyl1, ti, tf, env=-1):
        if cyl0 is None:
            check(_lib.lib().waves_set_design(self._h, env, 0, None, None, C.c_float(0), C.c_float(0)))
            return
        a, b = np.ascontiguousarray(cyl0, F32), np.ascontiguousarray(cyl1, F32)
        assert a.shape == b.shape and a.shape[1] == 4
        check(_lib.lib().waves_set_design(self._h, env, a.shape[0], a.ctypes.data_as(_lib.fp), b.ctypes.data_as(_lib.fp),
                                          C.c_float(ti), C.c_float(tf)))

    def set_design_batch(self, cyl0, cyl1, ti, tf):
        """One DesignInterpolator per environment over a common [ti, tf]: cyl0 / cyl1 are (n_env, ncyl, 4)."""
        a, b = np.ascontiguousarray(cyl0, F32), np.ascontiguousarray(cyl1, F32)
        assert a.shape == b.shape and a.shape[0] == self.n_env and a.shape[2] == 4
        check(_lib.lib().waves_set_design_batch(self._h, a.shape[1], a.ctypes.data_as(_lib.fp), b.ctypes.data_as(_lib.fp),
                                                C.c_float(ti), C.c_float(tf)))

    def set_speed_field(self, c, env=-1):
        check(_lib.lib().waves_set_speed_field(self._h, env, _ptr(c)))

    # dynamics ----------------------------------------------------------------
    def rhs(self, t, env=0):
        out = np.empty((12, self.ny, self.nx), dtype=F32)
        check(_lib.lib().waves_rhs(self._h, env, C.c_float(t), _ptr(out)))
        return out

    def step(self, t, mode=MODE_FUSED):
        check(_lib.lib().waves_step(self._h, C.c_float(t), mode))

    def integrate(self, tspan, mode=MODE_FUSED, energy=True, save_steps=(), frames=None, u_tot=None, u_inc=None):
        """(iter::Integrator)(ui, tspan, θ) + energy metric.  Returns (energy (n_env, steps+1, 3) | None, frames | None)."""
        ts = np.ascontiguousarray(tspan, F32)
        steps = len(ts) - 1
        en = np.empty((self.n_env, steps + 1, 3), dtype=F32) if energy is True else (energy if energy is not False else None)
        ss = np.ascontiguousarray(save_steps, np.int32)
        if len(ss) and frames is None:
            frames = np.empty((self.n_env, len(ss), 12, self.ny, self.nx), dtype=F32)
        check(_lib.lib().waves_integrate(self._h, ts.ctypes.data_as(_lib.fp), steps, mode, _ptr(en),
                                         ss.ctypes.data_as(_lib.ip) if len(ss) else None, len(ss), _ptr(frames),
                                         _ptr(u_tot), _ptr(u_inc)))
        return en, frames

    def set_traj_stride(self, stride: int):
        """u_tot / u_inc trajectories of integrate keep every `stride`-th frame: (n_env, steps // stride + 1, ny, nx)."""
        check(_lib.lib().waves_set_traj_stride(self._h, int(stride)))

    def set_graph(self, mode):
        """How integrate() issues kernels: 0 / False every kernel directly; 1 / True (default) CUDA graph replay for large
        batches, one launch per step for small ones; 2 additionally several steps per cooperative launch for small batches."""
        check(_lib.lib().waves_set_graph(self._h, int(mode)))

    def set_adjoint_checkpoint(self, every: int):
        """Reverse pass: checkpoint every `every` steps and re-run one segment at a time (0: automatic, from free memory)."""
        check(_lib.lib().waves_set_adjoint_checkpoint(self._h, int(every)))

    def adjoint(self, tspan, w_energy=None, dL_dzN=None, fwd_mode=MODE_FUSED, adj_mode=ADJ_EXACT, want_dc=True, out_dz0=None,
                out_dc=None, fused_reverse=True, march=True, ring=True):
        """rrule(::Integrator) + adjoint_sensitivity (src/dynamics.jl:97-128) from the current state.
        Returns (loss (n_env,), dL/dz0 (n_env,12,ny,nx), dL/dc (n_env,ny,nx) | None); out_dz0 / out_dc may be preallocated
        NumPy arrays or CUDA tensors (the gradients of a large batch are best left on the device)."""
        ts = np.ascontiguousarray(tspan, F32)
        steps = len(ts) - 1
        we = None if w_energy is None else np.ascontiguousarray(w_energy, F32)
        assert we is None or we.shape == (steps + 1, 3)
        an = None if dL_dzN is None else np.ascontiguousarray(dL_dzN, F32)
        gz = out_dz0 if out_dz0 is not None else np.empty((self.n_env, 12, self.ny, self.nx), dtype=F32)
        gc = out_dc if out_dc is not None else (np.empty((self.n_env, self.ny, self.nx), dtype=F32) if want_dc else None)
        loss = np.zeros(self.n_env, dtype=F32)
        if not fused_reverse:
            adj_mode |= ADJ_STAGEWISE   # per-stage reverse kernels even without dL/dc (cross-check of the fused reverse step)
        if ring is False:
            adj_mode |= ADJ_NO_RING     # the march kernel takes the interior only (cross-check of the PML-ring variants)
        elif ring == "always":
            adj_mode |= ADJ_RING        # ... and the PML ring on the march even for a small (latency-bound) batch
        if not march:
            adj_mode |= ADJ_TILES       # fused reverse step on the shared-memory tiles everywhere (cross-check of the march kernel)
        check(_lib.lib().waves_adjoint(self._h, ts.ctypes.data_as(_lib.fp), steps, fwd_mode, adj_mode, _ptr(we), _ptr(an),
                                       _ptr(gz), _ptr(gc), _ptr(loss)))
        return loss, gz, gc

    def observe(self, frames, resolution=(128, 128), out=None):
        """RLBase.state(env) image (src/env.jl:132-137): imresize(cat(U_tot of the saved frames, source shape), resolution).
        frames: (n_env, nsave, 12, ny, nx) as filled by integrate (NumPy or CUDA tensor) -> (n_env, nsave + 1, res_y, res_x)."""
        nsave = int(frames.shape[1])
        rx, ry = int(resolution[0]), int(resolution[1])
        if out is None:
            out = np.empty((self.n_env, nsave + 1, ry, rx), dtype=F32)
        check(_lib.lib().waves_observe(self._h, _ptr(frames), nsave, rx, ry, _ptr(out)))
        return out

    def energy(self, out=None):
        """(n_env, 3) energies of the current state; a CUDA tensor `out` is filled asynchronously on the engine's stream."""
        if out is None:
            out = np.empty((self.n_env, 3), dtype=F32)
        check(_lib.lib().waves_energy(self._h, _ptr(out)))
        return out

    def sync(self):
        check(_lib.lib().waves_sync(self._h))

    # slab plumbing -----------------------------------------------------------
    def halo_describe(self):
        d = _lib.HaloDesc()
        check(_lib.lib().waves_halo_describe(self._h, C.byref(d)))
        return d

    def halo_pack(self, lo, hi):
        check(_lib.lib().waves_halo_pack(self._h, _ptr(lo), _ptr(hi)))

    def halo_unpack(self, lo, hi):
        check(_lib.lib().waves_halo_unpack(self._h, _ptr(lo), _ptr(hi)))

    def peer_export(self) -> bytes:
        """This slab's waves_peer_info record (CUDA IPC handles + geometry) as bytes, to be shipped to the neighbouring ranks."""
        info = _lib.PeerInfo()
        check(_lib.lib().waves_peer_export(self._h, C.byref(info)))
        return bytes(info)

    def peer_attach(self, lo: bytes | None, hi: bytes | None):
        """Attach the neighbours' records: the fused step then stores its edge rows straight into their ghost rows."""
        a = _lib.PeerInfo.from_buffer_copy(lo) if lo is not None else None
        b = _lib.PeerInfo.from_buffer_copy(hi) if hi is not None else None
        check(_lib.lib().waves_peer_attach(self._h, C.byref(a) if a is not None else None, C.byref(b) if b is not None else None))

    def stream(self) -> int:
        return int(_lib.lib().waves_stream(self._h) or 0)

    # introspection -----------------------------------------------------------
    def launch_count(self) -> int:
        return int(_lib.lib().waves_launch_count(self._h))

    def graph_replays(self) -> int:
        return int(_lib.lib().waves_graph_replays(self._h))

    def profile(self, on: bool):
        check(_lib.lib().waves_profile(self._h, int(on)))

    def profile_read(self):
        ms, n = C.c_double(0), C.c_int64(0)
        check(_lib.lib().waves_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
