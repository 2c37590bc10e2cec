"""Mirror of the reference's call surface for the hot path: AcousticDynamics / Integrator / WaveEnv.

Same names, argument meaning and state as src/dynamics.jl:18-53,130-149 and src/env.jl:14-121; the
computation is one call into the CUDA library per `env(action)` / `iter(ui, tspan, θ)`.
"""
from __future__ import annotations

import numpy as np

from .engine import MODE_EXACT, MODE_FUSED, Engine
from .host import (WATER, DesignInterpolator, NoSource, TwoDim, build_action_space, build_gradient, build_pml,
                   build_tspan, get_dx, get_dy)

F32 = np.float32
FRAMESKIP = 10  # src/env.jl:90


class AcousticDynamics:
    """AcousticDynamics(dim, c0, pml_width, pml_scale) (src/dynamics.jl:130-149)."""

    def __init__(self, dim: TwoDim, c0, pml_width, pml_scale):
        self.dim, self.c0 = dim, F32(c0)
        self.pml_width, self.pml_scale = F32(pml_width), F32(pml_scale)
        self.grad = build_gradient(dim)               # rows of the sparse matrix
        self.pml = build_pml(dim, pml_width, pml_scale)  # 1-D profile


class Integrator:
    """Integrator(runge_kutta, dyn, dt) (src/dynamics.jl:18-53) bound to a device engine.

    θ is the pair the reference passes (src/env.jl:102): C is None (NoDesign / constant c0) or a
    DesignInterpolator; F is a Source / RandomPosGaussianSource / NoSource.
    """

    def __init__(self, dynamics: AcousticDynamics, dt, n_env=1, device=0, mode=MODE_FUSED):
        self.dynamics, self.dt, self.mode = dynamics, F32(dt), mode
        d = dynamics
        self.engine = Engine(d.dim.x, d.dim.y, d.c0, dt, d.pml_width, d.pml_scale, n_env=n_env, device=device,
                             sigma=d.pml, grad8=d.grad, d_omega=float(get_dx(d.dim) * get_dy(d.dim)))

    def build_tspan(self, ti, steps):
        return build_tspan(ti, self.dt, steps)

    def bind(self, theta, env=-1):
        C, Fsrc = theta
        if C is None:
            self.engine.set_design(None, None, 0, 0, env)
        else:
            self.engine.set_design(C.initial.table(), C.final.table(), C.ti, C.tf, env)
        self.engine.set_source(getattr(Fsrc, "shape", None), float(getattr(Fsrc, "freq", 0.0)), env)

    def __call__(self, ui, tspan, theta, energy=False):
        """Returns the (steps+1, 12, ny, nx) solution like the reference's cat(ui, ...; dims=4)."""
        self.bind(theta)
        self.engine.set_state(np.ascontiguousarray(ui, F32)[None] if np.ndim(ui) == 3 else ui)
        steps = len(tspan) - 1
        en, frames = self.engine.integrate(tspan, self.mode, energy=energy, save_steps=np.arange(steps + 1))
        return (frames[0], en[0]) if energy else frames[0]


class WaveEnv:
    """WaveEnv(dim; design_space, source, ...) (src/env.jl:14-67) with `env(action)` (src/env.jl:91-121)."""

    def __init__(self, dim: TwoDim, design_space=None, source=None, action_speed=250.0, c0=WATER, pml_width=2.0,
                 pml_scale=20000.0, dt=1e-5, integration_steps=100, actions=10, device=0, mode=MODE_FUSED, rng=None,
                 resolution=(128, 128)):
        self.dim, self.design_space = dim, design_space
        self.source = source if source is not None else NoSource()
        self.rng = rng or np.random.default_rng()
        self.action_speed, self.dt = F32(action_speed), F32(dt)
        self.integration_steps, self.actions = int(integration_steps), int(actions)
        self.resolution = tuple(resolution)
        assert all(n > r for n, r in zip(dim.size(), self.resolution)), "Resolution must be less than finite element grid."
        self.design = design_space.rand(self.rng) if design_space is not None else None
        self.iter = Integrator(AcousticDynamics(dim, c0, pml_width, pml_scale), dt, 1, device, mode)
        nx, ny = dim.size()
        # env.wave = the frames 80, 90, 100 of the last action, (nx,ny,12,3) in the reference, kept ON THE DEVICE like the
        # reference's gpu(env) does (src/env.jl:116); `env.wave` hands out a host copy on demand
        import torch
        self._wave_dev = torch.zeros((1, 3, 12, ny, nx), dtype=torch.float32, device=f"cuda:{device}")
        self._wave_host = None
        self.signal = np.zeros((self.integration_steps + 1, 3), dtype=F32)
        self.time_step = 0
        self.iter.engine.set_state(self._wave_dev[0, -1])
        self._source_bound = None

    @property
    def wave(self):
        """(3, 12, ny, nx) host copy of the kept frames (one device->host copy per action, made on first use)."""
        if self._wave_host is None:
            self._wave_host = self._wave_dev[0].cpu().numpy()
        return self._wave_host

    def time(self):  # src/env.jl:69-71
        return F32(F32(self.time_step) * self.dt)

    def build_tspan(self):  # src/env.jl:73-75
        return build_tspan(self.time(), self.dt, self.integration_steps)

    def is_terminated(self):  # src/env.jl:77-79
        return self.time_step >= self.actions * self.integration_steps

    def reset(self):  # src/env.jl:81-88
        self.time_step = 0
        self._wave_dev.zero_()
        self._wave_host = None
        self.design = self.design_space.rand(self.rng) if self.design_space is not None else None
        self.signal = np.zeros_like(self.signal)
        self.source.reset(self.rng)
        self.iter.engine.set_state(self._wave_dev[0, -1])
        self._source_bound = None

    def _bind_source(self):
        if self._source_bound is not self.source.shape:
            self.iter.engine.set_source(self.source.shape, float(self.source.freq))
            self._source_bound = self.source.shape

    def state(self):  # src/env.jl:132-137
        """(tspan, x, design) with x = imresize(cat(u_tot frames, source shape), resolution) as (4, res_y, res_x)."""
        self._bind_source()
        x = self.iter.engine.observe(self._wave_dev, self.resolution)[0]   # only (4, res_y, res_x) floats leave the GPU
        return self.build_tspan(), x, self.design

    def action_space(self):  # src/env.jl:143-145
        return build_action_space(self.design, F32(F32(self.action_speed * self.dt) * F32(self.integration_steps)))

    def reward(self):  # src/env.jl:147-149
        return F32(np.sum(self.signal, dtype=np.float64))

    def __call__(self, action, return_frames=False, frame_stride=1):
        """One environment step.  The device state and env.wave carry over between calls on the device, so only the design
        table, the tspan and the energy signal cross the host-device boundary (and the whole call replays one CUDA graph);
        return_frames adds the U trajectories a renderer wants (src/env.jl:120), every `frame_stride`-th frame."""
        eng = self.iter.engine
        tspan = self.build_tspan()
        ti = self.time()
        if self.design is None:
            nxt, interp = None, None
            eng.set_design(None, None, 0, 0)
        else:
            nxt = self.design_space(self.design, action)
            interp = DesignInterpolator(self.design, nxt, ti, tspan[-1])
            eng.set_design(self.design.table(), nxt.table(), ti, tspan[-1])
        self._bind_source()
        n = self.integration_steps
        ny, nx = self._wave_dev.shape[3:]
        nf = n // int(frame_stride) + 1
        u_tot = np.empty((1, nf, ny, nx), F32) if return_frames else None
        u_inc = np.empty((1, nf, ny, nx), F32) if return_frames else None
        eng.set_traj_stride(frame_stride if return_frames else 1)
        en, _ = eng.integrate(tspan, self.iter.mode, energy=True, save_steps=[n - 2 * FRAMESKIP, n - FRAMESKIP, n],
                              frames=self._wave_dev, u_tot=u_tot, u_inc=u_inc)
        self.signal = en[0]
        self.design = nxt
        self._wave_host = None
        self.time_step += n
        return tspan, interp, (u_tot[0] if return_frames else None), (u_inc[0] if return_frames else None)
