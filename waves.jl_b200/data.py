"""Mirror of the reference's episode layer (src/data.jl, src/env.jl:151-157) on top of the device engine, plus the batched
form the reference lacks: many WaveEnv stepping in lockstep on one handle.

Only bookkeeping lives here; the integration, the energy signal and the observation image run in the CUDA library.
Episodes are saved as .npz or, with `Episode.save_bson`, in the reference's own BSON layout (src/data.jl:60-71; see bson.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .engine import MODE_FUSED, Engine
from .env import FRAMESKIP, WaveEnv
from .host import WATER, DesignInterpolator, TwoDim, build_action_space, build_tspan

F32 = np.float32


def flatten_repeated_last_dim(x: np.ndarray) -> np.ndarray:
    """src/utils.jl:20-31 for the C-order image of the reference's arrays: x is (n_windows, n_samples[, k]) (the reference's
    trailing dimension is our leading one); consecutive windows share their boundary sample, which is kept once."""
    x = np.asarray(x)
    return np.concatenate([x[0]] + [w[1:] for w in x[1:]], axis=0)


class RandomDesignPolicy:
    """src/env.jl:151-157."""

    def __init__(self, a_space, rng=None):
        self.a_space, self.rng = a_space, rng or np.random.default_rng()

    def __call__(self, env=None):
        return self.a_space.rand(self.rng)


@dataclass
class Episode:
    """src/data.jl:3-8: states, actions, tspans and energy signals of one episode."""
    s: list = field(default_factory=list)   # (tspan, image (4, res, res), design) per action  (WaveEnvState, src/env.jl:5-10)
    a: list = field(default_factory=list)
    t: list = field(default_factory=list)   # (steps+1,) per action
    y: list = field(default_factory=list)   # (steps+1, 3) per action

    def __len__(self):
        return len(self.s)

    def save(self, path: str):
        np.savez_compressed(path, images=np.stack([s[1] for s in self.s]), t=np.stack(self.t), y=np.stack(self.y),
                            designs=np.stack([s[2].table() for s in self.s]), actions=np.stack([a.table() for a in self.a]))

    def save_bson(self, path: str, dim, scatterers: str = "AdjustableRadiiScatterers"):
        """FileIO.save(episode, path) (src/data.jl:60-62): the (s, a, t, y) document `Episode(path = ...)` loads."""
        from . import bson
        bson.save_episode(self, dim, path, scatterers)


def generate_episode(policy, env: WaveEnv, reset: bool = True) -> Episode:
    """generate_episode!(policy, env) (src/data.jl:12-33)."""
    ep = Episode()
    if reset:
        env.reset()
    while not env.is_terminated():
        ep.s.append(env.state())
        action = policy(env)
        ep.a.append(action)
        ep.t.append(env.build_tspan())
        env(action)
        ep.y.append(env.signal.copy())
    return ep


def prepare_data(ep: Episode, horizon: int):
    """prepare_data(ep, horizon) (src/data.jl:35-58): sliding windows of `horizon` consecutive actions."""
    s, a, t, y = [], [], [], []
    n = horizon - 1
    for i in range(len(ep) - n):
        s.append(ep.s[i])
        a.append(ep.a[i:i + horizon])
        t.append(flatten_repeated_last_dim(np.stack(ep.t[i:i + horizon])))
        y.append(flatten_repeated_last_dim(np.stack(ep.y[i:i + horizon])))
    return s, a, t, y


class BatchWaveEnv:
    """n_env WaveEnv (src/env.jl:14-121) that step in lockstep on ONE device handle: every environment has its own design,
    source and wavefield, shares dim / dt / integration_steps / actions, and `env(actions)` advances all of them with one
    fused launch set per RK4 step.  Environment e of the batch is bitwise the WaveEnv built from the same inputs."""

    def __init__(self, dim: TwoDim, design_space, sources, action_speed=250.0, c0=WATER, pml_width=2.0, pml_scale=20000.0, dt=1e-5,
                 integration_steps=100, actions=10, device=0, mode=MODE_FUSED, rngs=None, resolution=(128, 128)):
        self.dim, self.design_space, self.sources = dim, design_space, list(sources)
        self.n_env = len(self.sources)
        self.rngs = rngs or [np.random.default_rng(e) for e in range(self.n_env)]
        self.action_speed, self.dt = F32(action_speed), F32(dt)
        self.integration_steps, self.actions, self.mode = int(integration_steps), int(actions), mode
        self.resolution = tuple(resolution)
        self.engine = Engine(dim.x, dim.y, c0, dt, pml_width, pml_scale, n_env=self.n_env, device=device)
        nx, ny = dim.size()
        self.designs = [design_space.rand(r) for r in self.rngs]
        # env.wave of every environment stays on the device (src/env.jl:116 under gpu(env)); `.wave` copies it out on demand
        import torch
        self._wave_dev = torch.zeros((self.n_env, 3, 12, ny, nx), dtype=torch.float32, device=f"cuda:{device}")
        self._wave_host = None
        self.signal = np.zeros((self.n_env, self.integration_steps + 1, 3), dtype=F32)
        self.time_step = 0
        self.engine.set_state(self._wave_dev[:, -1].contiguous())
        for e, src in enumerate(self.sources):
            self.engine.set_source(src.shape, float(src.freq), env=e)

    @property
    def wave(self):
        """(n_env, 3, 12, ny, nx) host copy of the kept frames, made on first use after an action."""
        if self._wave_host is None:
            self._wave_host = self._wave_dev.cpu().numpy()
        return self._wave_host

    def time(self):
        return F32(F32(self.time_step) * self.dt)

    def build_tspan(self):
        return build_tspan(self.time(), self.dt, self.integration_steps)

    def is_terminated(self):
        return self.time_step >= self.actions * self.integration_steps

    def action_spaces(self):
        scale = F32(F32(self.action_speed * self.dt) * F32(self.integration_steps))
        return [build_action_space(d, scale) for d in self.designs]

    def state(self):
        """Per environment (tspan, image (4, res_y, res_x), design): RLBase.state (src/env.jl:132-137)."""
        x = self.engine.observe(self._wave_dev, self.resolution)   # device frames in, (n_env, 4, res_y, res_x) floats out
        ts = self.build_tspan()
        return [(ts, x[e], self.designs[e]) for e in range(self.n_env)]

    def reward(self):
        return np.sum(self.signal, axis=(1, 2), dtype=np.float64).astype(F32)

    def __call__(self, actions):
        """One env(action) (src/env.jl:91-121) for every environment."""
        tspan, ti = self.build_tspan(), self.time()
        nxt = [self.design_space(d, a) for d, a in zip(self.designs, actions)]
        self.engine.set_design_batch(np.stack([d.table() for d in self.designs]), np.stack([d.table() for d in nxt]), ti, tspan[-1])
        n = self.integration_steps
        self.signal, _ = self.engine.integrate(tspan, self.mode, energy=True, save_steps=[n - 2 * FRAMESKIP, n - FRAMESKIP, n],
                                               frames=self._wave_dev)
        self._wave_host = None
        interps = [DesignInterpolator(d, x, ti, tspan[-1]) for d, x in zip(self.designs, nxt)]
        self.designs = nxt
        self.time_step += n
        return tspan, interps


def generate_episodes(policies, env: BatchWaveEnv):
    """generate_episode! for every environment of a batch at once (the dataset loop of scripts/data.jl): one Episode each."""
    eps = [Episode() for _ in range(env.n_env)]
    while not env.is_terminated():
        states = env.state()
        actions = [p(env) for p in policies]
        ts = env.build_tspan()
        env(actions)
        for e, ep in enumerate(eps):
            ep.s.append(states[e])
            ep.a.append(actions[e])
            ep.t.append(ts)
            ep.y.append(env.signal[e].copy())
    return eps
