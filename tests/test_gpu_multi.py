"""Multi-GPU tests (need >= 2 B200s; skipped otherwise): slab decomposition with NCCL halo exchange must
reproduce the single-GPU run bit for bit, and env-sharded batches need no communication."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import waves_b200 as wb
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for n, steps, gs in ((512, 12, 11.0), (257, 9, 5.5)):   # 257: odd row count, pitch != nx
    dim = wb.TwoDim(gs, n)
    rng = np.random.default_rng(n)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(np.float32)
    shape = wb.build_normal(dim, [[-0.3 * gs, 0.04 * gs]], [0.03 * gs], [1.0])
    ts = wb.build_tspan(0.0, 1e-5, steps)
    slab = wb.SlabEngine(dim.x, dim.y, wb.WATER, 1e-5, 0.2 * gs, 20000.0, device=local)
    pslab = wb.SlabEngine(dim.x, dim.y, wb.WATER, 1e-5, 0.2 * gs, 20000.0, device=local, peer=True)   # NVLink peer stores
    slab.set_source_global(shape, 1000.0)
    pslab.set_source_global(shape, 1000.0)
    for mode in (wb.MODE_FUSED, wb.MODE_EXACT, "peer"):
        if mode == "peer":
            pslab.set_state_global(u0)
            en = pslab.integrate(ts, wb.MODE_FUSED)
            pslab.set_state_global(u0)                      # a second run on the same handles (flags keep counting)
            en = pslab.integrate(ts, wb.MODE_FUSED)
            full = pslab.gather_state()
            mode = wb.MODE_FUSED
            key = (n, "peer")
        else:
            slab.set_state_global(u0)
            en = slab.integrate(ts, mode)
            full = slab.gather_state()
            key = (n, mode)
        if rank == 0:
            eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 0.2 * gs, 20000.0, device=0)
            eng.set_state(u0[None]); eng.set_source(shape, 1000.0)
            ren, _ = eng.integrate(ts, mode)
            ref = eng.get_state(0)
            out[key] = (bool(np.array_equal(full, ref)), float(np.abs(full - ref).max()), float(np.abs(en - ren[0]).max() / ren[0].max()))
            eng.close()
    slab.close()
    pslab.close()
if rank == 0:
    print("RESULT", out)
dist.barrier(); dist.destroy_process_group()
'''


def test_slab_decomposition_matches_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    script = tmp_path / "slab_worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1]
    res = eval(line[len("RESULT "):])
    print(res)
    for mode, (bitwise, maxd, erel) in res.items():
        assert bitwise, f"mode {mode}: slab run differs from single GPU (max |d| = {maxd}): {res}"
        assert erel < 1e-6
