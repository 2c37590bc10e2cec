"""world_size-2 gloo tests (CPU) of the multi-rank host logic: environment sharding, slab partitioning and
the neighbour halo exchange pattern (the CUDA pack/unpack kernels are replaced by NumPy stand-ins with the
same contract; the real ones are covered by tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import waves_b200 as wb

HALO = 4


def test_shard_envs_partitions_exactly():
    for total, world in [(1024, 8), (10, 4), (3, 2), (7, 7)]:
        got = [list(wb.shard_envs(total, r, world)) for r in range(world)]
        assert sum(got, []) == list(range(total))
        assert max(map(len, got)) - min(map(len, got)) <= 1


def test_slab_rows_cover_grid():
    for ny, world in [(16384, 8), (700, 2), (701, 4)]:
        rows = [wb.slab_rows(ny, r, world) for r in range(world)]
        assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
        for (a, n), (b, _) in zip(rows, rows[1:]):
            assert a + n == b
    with pytest.raises(ValueError):
        wb.slab_rows(20, 0, 4)


class FakeSlab:
    """CPU stand-in with the halo contract of the engine: state (planes, rows + ghosts, nx)."""

    def __init__(self, full, row0, n, ny):
        self.gtop, self.gbot = (HALO if row0 > 0 else 0), (HALO if row0 + n < ny else 0)
        self.u = np.full((full.shape[0], n + self.gtop + self.gbot, full.shape[2]), np.nan, np.float32)
        self.u[:, self.gtop:self.gtop + n] = full[:, row0:row0 + n]
        self.n = n

    def pack(self, lo, hi):
        if lo is not None:
            lo.copy_(torch.from_numpy(self.u[:, self.gtop:self.gtop + HALO].reshape(-1).copy()))
        if hi is not None:
            hi.copy_(torch.from_numpy(self.u[:, self.gtop + self.n - HALO:self.gtop + self.n].reshape(-1).copy()))

    def unpack(self, lo, hi):
        if lo is not None:
            self.u[:, :HALO] = lo.numpy().reshape(self.u.shape[0], HALO, -1)
        if hi is not None:
            self.u[:, self.gtop + self.n:] = hi.numpy().reshape(self.u.shape[0], HALO, -1)


def _worker(rank, world, port, ny, nx, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = np.arange(3 * ny * nx, dtype=np.float32).reshape(3, ny, nx)
        row0, n = wb.slab_rows(ny, rank, world)
        slab = FakeSlab(full, row0, n, ny)
        ex = wb.HaloExchanger(rank, world, 3 * HALO * nx, "cpu")
        ex.exchange(slab.pack, slab.unpack)
        lo, hi = row0 - slab.gtop, row0 + n + slab.gbot
        ok = np.array_equal(slab.u, full[:, lo:hi])
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        envs = list(wb.shard_envs(5, rank, world))
        q.put((rank, bool(ok), float(t.item()), envs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 40, 8, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _, _ in res), "ghost rows must equal the neighbour's boundary rows"
    assert all(t == float(world) for _, _, t, _ in res)
    assert sum((e for _, _, _, e in res), []) == list(range(5))
