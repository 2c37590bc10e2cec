"""BASELINE configs[3]: ONE large grid with PML, slab-decomposed along y over N GPUs.  Two halo transports: an NCCL send/recv
exchange after every fused RK4 step, or (`peer`) the edge rows stored straight into the neighbours' ghost rows over NVLink by
the step kernel itself.  `python scripts/bench_slab.py [n] [steps] [peer]` (N = 1) or under torchrun; bench.py imports
`measure` for its "slab" sub-record.
Domain: TwoDim(n * 30/699 / 2, n) keeps the reference's dx (TwoDim(15, 16384) is CFL-unstable)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waves_b200 as wb  # noqa: E402


def real_bytes_per_cell_update():
    """DRAM bytes one cell-update of this workload really moves (ncu, committed under profiles/): fewer than the 96 B the
    metric counts, because fields that are constant where sigma is zero are not re-copied."""
    p = os.path.join(ROOT, "profiles", "r2_slab_traffic.json")
    if os.path.exists(p):
        return float(json.load(open(p))["dram_bytes_per_cell_update"]), "ncu, committed (profiles/r2_slab_traffic.json)"
    return 60.0, "estimate: lean interior reads U, Vx, Vy, P (+ U_inc) and writes U, Vx, Vy = 32 + 28 B per cell of both wavefields"


def _source_rows(dim, lo, hi):
    # only this rank's rows of the Gaussian source are ever built (the full plane is 1 GB at 16384^2)
    xx, yy = dim.x[None, :], dim.y[lo:hi][:, None]
    return (np.float32(1.0 / (2 * np.pi * 0.3 ** 2)) * np.exp(-((xx + 10.0) ** 2 + yy ** 2) / np.float32(2 * 0.3 ** 2))).astype(np.float32)


def measure(n=16384, steps=20, peer=True, rank=0, world=1, local=0, check_single=True):
    """All ranks call this (process group initialised when world > 1).  Returns the record on rank 0, None elsewhere."""
    import torch.distributed as dist
    gs = np.float32(0.5 * (n - 1) * 30.0 / 699.0)
    dim = wb.TwoDim(gs, n)
    ts = wb.build_tspan(0.0, 1e-5, steps)
    row0, ny = wb.slab_rows(n, rank, world) if world > 1 else (0, n)
    lo = row0 - (4 if rank > 0 else 0)
    hi = row0 + ny + (4 if rank < world - 1 else 0)
    slab = None
    if world > 1:
        slab = wb.SlabEngine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, device=local, peer=peer)
        eng = slab.engine
    else:
        eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, device=local)
    eng.set_source(_source_rows(dim, lo, hi), 1000.0)
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)
    step_no = [0]

    def run(k):
        for _ in range(k):
            eng.step(float(ts[step_no[0] % steps]), wb.MODE_FUSED | wb.STEP_ASYNC)
            step_no[0] += 1
            if world > 1 and not peer:
                slab.exchange()

    run(3)   # warm-up (3 steps; the timed steps continue from there, the check below repeats the same 3 + steps sequence)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run(steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=f"cuda:{local}")
    et = torch.tensor(eng.energy()[0].astype(np.float64), device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(et)
    out = None
    if rank == 0:
        v = n * n * steps / (ms.item() * 1e-3) / 1e9
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
        bpc, src = real_bytes_per_cell_update()
        out = {"workload": f"single {n}^2 grid with PML (BASELINE configs[3]), slab-decomposed over {world} B200, "
                           + ("edge rows stored into the neighbours' ghost rows over NVLink peer memory by the step kernel"
                              if (peer and world > 1) else ("NCCL halo exchange every RK4 step" if world > 1 else "one handle")),
               "n": n, "n_gpus": world, "steps": steps, "ms_per_step": round(ms.item() / steps, 4), "value": round(v, 2),
               "unit": "Gcell-updates/s", "real_dram_bytes_per_cell_update": round(bpc, 1), "real_dram_source": src,
               "real_dram_gbs": round(v * bpc / world, 1), "frac_real": round(v * bpc / (peak * world), 4),
               "frac_at_96B_contract": round(v * 96 / (peak * world), 4),
               "energy_tot_inc_sc": [float(x) for x in et.cpu().numpy()]}
    if world > 1 and check_single:
        # the same 3 + steps RK4 steps on ONE handle (rank 0; 26 GB): the slab run must give the same energies
        if rank == 0:
            ref = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, device=local)
            ref.set_source(_source_rows(dim, 0, n), 1000.0)
            for i in range(3 + steps):
                ref.step(float(ts[i % steps]), wb.MODE_FUSED | wb.STEP_ASYNC)
            e1 = ref.energy()[0].astype(np.float64)
            ref.close()
            e8 = np.array(out["energy_tot_inc_sc"])
            out["energy_rel_diff_vs_single_gpu"] = float(np.abs(e8 - e1).max() / max(e1.max(), 1e-300))
        dist.barrier()
    eng.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    peer = len(sys.argv) > 3 and sys.argv[3] == "peer"
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rec = measure(n, steps, peer, rank, world, local)
    if rec is not None:
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
