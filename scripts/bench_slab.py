"""BASELINE configs[3]: ONE large grid with PML, slab-decomposed along y over N GPUs with a halo exchange (NCCL send/recv
over NVLink) after every fused RK4 step.  `python scripts/bench_slab.py [n] [steps]` (N = 1) or under torchrun.
Domain: TwoDim(n * 30/699 / 2, n) keeps the reference's dx (TwoDim(15, 16384) is CFL-unstable)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
peer = len(sys.argv) > 3 and sys.argv[3] == "peer"   # NVLink peer stores instead of the NCCL halo exchange
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gs = np.float32(0.5 * (n - 1) * 30.0 / 699.0)
dim = wb.TwoDim(gs, n)
ts = wb.build_tspan(0.0, 1e-5, steps)
row0, ny = wb.slab_rows(n, rank, world) if world > 1 else (0, n)
lo = row0 - (4 if rank > 0 else 0)
hi = row0 + ny + (4 if rank < world - 1 else 0)
# only this rank's rows of the Gaussian source are ever built (the full plane is 1 GB at 16384^2)
sub = wb.TwoDim(gs, n)
sub_y = dim.y[lo:hi]
xx, yy = dim.x[None, :], sub_y[:, None]
shape = (np.float32(1.0 / (2 * np.pi * 0.3 ** 2)) * np.exp(-((xx + 10.0) ** 2 + yy ** 2) / np.float32(2 * 0.3 ** 2))).astype(np.float32)
if world > 1:
    slab = wb.SlabEngine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, device=local, peer=peer)
    eng = slab.engine
else:
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, device=local)
eng.set_source(shape, 1000.0)
stream = torch.cuda.ExternalStream(eng.stream(), device=local)


def run(k):
    for i in range(k):
        eng.step(float(ts[i % steps]), wb.MODE_FUSED | wb.STEP_ASYNC)
        if world > 1 and not peer:
            slab.exchange()


run(3)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
run(steps)
ev1.record(stream)
torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
e = eng.energy()[0].astype(np.float64)
et = torch.tensor(e, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(et)
if rank == 0:
    v = n * n * steps / (ms.item() * 1e-3) / 1e9
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({"workload": f"single {n}^2 grid with PML, slab-decomposed over {world} B200, " + ("edge rows stored into the neighbours' ghost rows over NVLink peer memory" if peer else "NCCL halo exchange every RK4 step"),
                      "n_gpus": world, "steps": steps, "ms_per_step": round(ms.item() / steps, 3), "value": round(v, 2),
                      "unit": "Gcell-updates/s", "frac_of_hbm_roofline": round(v * 96 / (peak * world), 4),
                      "energy_tot_inc_sc": [float(x) for x in et.cpu().numpy()]}))
eng.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
