"""Multi-GPU plumbing: one process per GPU (torch.distributed), two partitionings.

 * batches of independent environments: contiguous blocks of environments per rank, NO data-path collective;
 * one large grid: 1-D slab decomposition along y (dim 2, the strided dimension, so every slab stays
   x-contiguous) with a nearest-neighbour exchange of WAVES_HALO = 4 rows x 12 fields after every RK4 step
   (one ghost row per RK stage; the fused kernel recomputes the 4-row halo instead of exchanging per stage).

torch.distributed is only the transport (NCCL over NVLink on GPUs, gloo in the CPU tests); the compute is
the CUDA library.
"""
from __future__ import annotations

import numpy as np

HALO = 4  # WAVES_HALO in include/waves_b200.h


def shard_envs(n_env_total: int, rank: int, world: int):
    """Contiguous block of environment indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_env_total), int(world))
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def slab_rows(ny: int, rank: int, world: int):
    """(row0, n_rows) of the slab of a grid with `ny` rows owned by `rank`; every slab needs >= 2*HALO rows."""
    base, rem = divmod(int(ny), int(world))
    if base < 2 * HALO:
        raise ValueError(f"{ny} rows over {world} ranks leaves slabs thinner than {2 * HALO} rows")
    row0 = rank * base + min(rank, rem)
    return row0, base + (1 if rank < rem else 0)


class HaloExchanger:
    """Exchanges the boundary rows of neighbouring slabs.

    `pack(lo, hi)` / `unpack(lo, hi)` are the engine's halo kernels (waves_halo_pack / waves_halo_unpack);
    `lo`/`hi` are torch tensors living wherever the transport wants them (CUDA for NCCL, CPU for gloo).
    Rank r sends its first owned rows to r-1 (they become r-1's upper ghost rows) and its last owned rows
    to r+1, non-periodically.
    """

    def __init__(self, rank: int, world: int, n_floats: int, device, group=None):
        import torch
        self.rank, self.world, self.group = rank, world, group
        self.has_lo, self.has_hi = rank > 0, rank < world - 1
        mk = lambda on: torch.empty(n_floats if on else 0, dtype=torch.float32, device=device)
        self.send_lo, self.recv_lo = mk(self.has_lo), mk(self.has_lo)
        self.send_hi, self.recv_hi = mk(self.has_hi), mk(self.has_hi)

    def exchange(self, pack, unpack):
        import torch.distributed as dist
        pack(self.send_lo if self.has_lo else None, self.send_hi if self.has_hi else None)
        ops = []
        if self.has_lo:
            ops += [dist.P2POp(dist.isend, self.send_lo, self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, self.recv_lo, self.rank - 1, self.group)]
        if self.has_hi:
            ops += [dist.P2POp(dist.isend, self.send_hi, self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, self.recv_hi, self.rank + 1, self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        unpack(self.recv_lo if self.has_lo else None, self.recv_hi if self.has_hi else None)


class SlabEngine:
    """One rank's slab of a single large grid (BASELINE config 4).  Requires an initialised process group."""

    def __init__(self, x, y, c0, dt, pml_width, pml_scale, device, rank=None, world=None, peer=False):
        import torch
        import torch.distributed as dist

        from .engine import Engine
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.row0, self.ny = slab_rows(len(y), self.rank, self.world)
        self.engine = Engine(x, y, c0, dt, pml_width, pml_scale, n_env=1, device=device, ny_local=self.ny, row0=self.row0)
        d = self.engine.halo_describe()
        self.stream = torch.cuda.ExternalStream(self.engine.stream(), device=device)
        self.halo = HaloExchanger(self.rank, self.world, d.n_planes * d.block_floats, torch.device("cuda", device))
        self._torch = torch
        # peer mode: the fused step stores its edge rows straight into the neighbours' ghost rows over NVLink (CUDA IPC
        # mappings + stream-ordered flags); the NCCL exchange is only used once after a state was set
        self.peer = bool(peer) and self.world > 1
        if self.peer:
            infos = [None] * self.world
            dist.all_gather_object(infos, self.engine.peer_export())
            self.engine.peer_attach(infos[self.rank - 1] if self.rank > 0 else None,
                                    infos[self.rank + 1] if self.rank < self.world - 1 else None)
            dist.barrier()

    def exchange(self):
        with self._torch.cuda.stream(self.stream):   # NCCL ordered with the engine's own stream
            self.halo.exchange(self.engine.halo_pack, self.engine.halo_unpack)

    def set_state_global(self, u12_global):
        """Scatter: every rank takes its rows of a (12, ny_global, nx) array, then ghost rows are exchanged."""
        import torch.distributed as dist
        if self.peer:
            self.engine.sync()
            dist.barrier()     # nobody may still be storing into this rank's ghost rows
        self.engine.set_state(np.ascontiguousarray(u12_global[:, self.row0:self.row0 + self.ny])[None])
        self.exchange()
        if self.peer:
            self.engine.sync()
            dist.barrier()

    def set_source_global(self, shape_global, freq):
        """Every rank takes its rows of the (ny_global, nx) shape INCLUDING the ghost rows of its slab."""
        if shape_global is None:
            self.engine.set_source(None, freq)
            return
        lo = self.row0 - (HALO if self.rank > 0 else 0)
        hi = self.row0 + self.ny + (HALO if self.rank < self.world - 1 else 0)
        self.engine.set_source(np.ascontiguousarray(shape_global[lo:hi]), freq)

    def integrate(self, tspan, mode=0, energy=True):
        """RK4 steps with one halo exchange per step; returns the global (steps+1, 3) energy trace (all ranks)."""
        import torch.distributed as dist
        torch = self._torch
        steps = len(tspan) - 1
        from .engine import MODE_FUSED, STEP_ASYNC
        peer = self.peer and mode == MODE_FUSED
        # per-step energies stay on the device (written on the engine's stream): one all-reduce at the end, no host sync per step
        d_en = torch.zeros((steps + 1, 3), dtype=torch.float32, device=f"cuda:{self.engine.device}") if energy else None
        if energy:
            self.engine.energy(out=d_en[0:1])
        for n in range(steps):
            self.engine.step(float(tspan[n]), mode | STEP_ASYNC)   # the exchange is ordered on the engine's stream
            if not peer:
                self.exchange()
            if energy:
                self.engine.energy(out=d_en[n + 1:n + 2])
        self.engine.sync()
        if not energy:
            return None
        t = d_en.double()
        dist.all_reduce(t)
        return t.cpu().numpy().astype(np.float32)

    def gather_state(self):
        """All ranks receive the full (12, ny_global, nx) state (test/diagnostic helper)."""
        import torch.distributed as dist
        local = self.engine.get_state(0)
        parts = [None] * self.world
        dist.all_gather_object(parts, local)
        return np.concatenate(parts, axis=1)

    def close(self):
        self.engine.close()


class LocalSlabGroup:
    """All slabs of one large grid driven by ONE process: `world` slab handles on the given devices (repeat a device ordinal to
    put several slabs on one GPU), ghost rows moved with waves_halo_pack -> device copy -> waves_halo_unpack.  No process group.
    This is the single-process form of SlabEngine (same slab geometry, same kernels), and what lets the slab path be checked
    bit for bit against the single-handle run on a box with one GPU."""

    def __init__(self, x, y, c0, dt, pml_width, pml_scale, devices):
        import torch

        from .engine import Engine
        self._torch, self.world = torch, len(devices)
        self.devices = [int(d) for d in devices]
        self.rows = [slab_rows(len(y), r, self.world) for r in range(self.world)]
        self.engines = [Engine(x, y, c0, dt, pml_width, pml_scale, n_env=1, device=d, ny_local=ny, row0=row0)
                        for d, (row0, ny) in zip(self.devices, self.rows)]
        self.buf = []
        for r, eng in enumerate(self.engines):
            d = eng.halo_describe()
            mk = lambda on: torch.empty(d.n_planes * d.block_floats if on else 0, dtype=torch.float32, device=f"cuda:{self.devices[r]}")
            # send_lo, send_hi, recv_lo, recv_hi
            self.buf.append([mk(r > 0), mk(r < self.world - 1), mk(r > 0), mk(r < self.world - 1)])

    def sync(self):
        for eng in self.engines:
            eng.sync()

    def exchange(self):
        W = self.world
        for r, eng in enumerate(self.engines):
            eng.halo_pack(self.buf[r][0] if r > 0 else None, self.buf[r][1] if r < W - 1 else None)
        self.sync()
        for r in range(W):  # my first rows -> the lower neighbour's upper ghost rows, my last rows -> the upper neighbour's lower ones
            if r > 0:
                self.buf[r - 1][3].copy_(self.buf[r][0])
            if r < W - 1:
                self.buf[r + 1][2].copy_(self.buf[r][1])
        self._torch.cuda.synchronize()
        for r, eng in enumerate(self.engines):
            eng.halo_unpack(self.buf[r][2] if r > 0 else None, self.buf[r][3] if r < W - 1 else None)
        self.sync()

    def set_state_global(self, u12_global):
        for eng, (row0, ny) in zip(self.engines, self.rows):
            eng.set_state(np.ascontiguousarray(u12_global[:, row0:row0 + ny])[None])
        self.exchange()

    def set_source_global(self, shape_global, freq):
        for r, (eng, (row0, ny)) in enumerate(zip(self.engines, self.rows)):
            lo = row0 - (HALO if r > 0 else 0)
            hi = row0 + ny + (HALO if r < self.world - 1 else 0)
            eng.set_source(None if shape_global is None else np.ascontiguousarray(shape_global[lo:hi]), freq)

    def integrate(self, tspan, mode=0):
        """RK4 steps with one exchange per step -> global (steps+1, 3) energy trace (float64 sum of the slabs' energies)."""
        from .engine import STEP_ASYNC
        steps = len(tspan) - 1
        en = np.zeros((steps + 1, 3), dtype=np.float64)
        en[0] = sum(eng.energy()[0].astype(np.float64) for eng in self.engines)
        for n in range(steps):
            for eng in self.engines:
                eng.step(float(tspan[n]), mode | STEP_ASYNC)
            self.exchange()
            en[n + 1] = sum(eng.energy()[0].astype(np.float64) for eng in self.engines)
        return en.astype(np.float32)

    def gather_state(self):
        return np.concatenate([eng.get_state(0) for eng in self.engines], axis=1)

    def close(self):
        for eng in self.engines:
            eng.close()
