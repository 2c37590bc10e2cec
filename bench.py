#!/usr/bin/env python
"""bench.py -- fused 2D PML RK4 wave step: Gcell-updates/s on 1/2/4/8 B200 vs the HBM roofline.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): a batch of independent
700^2 WaveEnv environments (triple-ring design + Gaussian source), 128 per GPU (1024 over 8 GPUs, weak
scaling, no communication).  One bench "step" = one `env(action)` for every environment of the batch:
100 fused RK4 steps + the 101x3 energy trace (src/env.jl:91-121).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--rk4-steps S]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)
  python bench.py --impl reference ...                      (reference CPU path: the oracle port, all host cores)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 700
GRID_SIZE = 15.0
ALG_BYTES_PER_CELL_UPDATE = 96  # 12 fp32 read + 12 fp32 written per cell per RK4 step
METRIC = "Gcell-updates/s of fused 2D PML RK4 step (1/2/4/8 B200), % of HBM roofline"
UNIT = "Gcell-updates/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_env_inputs(global_env):
    """Deterministic per-environment synthetic inputs (config 2 seeds: NumPy default_rng(seed=env))."""
    import waves_b200 as wb
    rng = np.random.default_rng(global_env)
    ds = make_env_inputs.ds
    d0 = ds.rand(rng)
    acts = [wb.build_action_space(d0, 0.25).rand(rng) for _ in range(2)]
    d1 = ds(d0, acts[0])
    mu_y = float(rng.uniform(-10.0, 10.0))
    return d0, d1, mu_y


def run_ours(args):
    import torch
    import waves_b200 as wb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    E, S, n = args.envs_per_gpu, args.rk4_steps, N_GRID
    dim = wb.TwoDim(GRID_SIZE, n)
    make_env_inputs.ds = wb.build_triple_ring_design_space()
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=E, device=local)
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)

    # synthetic inputs: per-env design pair + source position; state starts at rest like reset!(env)
    designs = []
    for e in range(E):
        d0, d1, mu_y = make_env_inputs(rank * E + e)
        designs.append((d0.table(), d1.table()))
        eng.set_source(wb.build_normal(dim, [[-10.0, mu_y]], [0.3], [1.0]), 1000.0, env=e)
    # pinned host staging for the per-step inputs (design tables) and outputs (energy signal)
    ncyl = designs[0][0].shape[0]
    h_design = torch.empty((E, 2, ncyl, 4), dtype=torch.float32).pin_memory()
    for e in range(E):
        h_design[e, 0] = torch.from_numpy(designs[e][0])
        h_design[e, 1] = torch.from_numpy(designs[e][1])
    h_energy = torch.empty((E, S + 1, 3), dtype=torch.float32).pin_memory()
    d_frames = torch.empty((E, 3, 12, n, n), dtype=torch.float32, device=f"cuda:{local}")  # env.wave stays on the device
    save = np.array([S - 20, S - 10, S], dtype=np.int32) if S >= 20 else np.array([S], dtype=np.int32)
    hd = h_design.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    time_step = [0]

    def bind_designs(ts):
        eng.set_design_batch(hd[:, 0], hd[:, 1], ts[0], ts[-1])

    def one_step(e2e):
        """One env(action) for the whole batch.  e2e: host design tables in, energy signal out, frames kept on device."""
        ts = wb.build_tspan(np.float32(np.float32(time_step[0]) * np.float32(1e-5)), 1e-5, S)
        if e2e:
            bind_designs(ts)
            eng.integrate(ts, wb.MODE_FUSED, energy=h_energy.numpy(), save_steps=save, frames=d_frames[:, :len(save)])
        else:
            eng.integrate(ts, wb.MODE_FUSED, energy=h_energy.numpy())
        time_step[0] += S

    def timed(k, e2e):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(k):
            one_step(e2e)
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        if e2e:
            ms = max(ms, 0.0)
            ms = wall * 1e3  # end to end is a host-visible latency: wall clock between the barriers
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.launch_count() - l0

    ts0 = wb.build_tspan(0.0, 1e-5, S)
    bind_designs(ts0)
    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(1):
        one_step(True)
    ms_e2e, _ = timed(args.steps, True)

    # dominant-kernel duration, measured live with CUDA events on the launching stream (separate pass)
    eng.profile(True)
    one_step(False)
    fused_ms, fused_n = eng.profile_read()
    eng.profile(False)

    cells_step = world * E * n * n * S
    value = cells_step * args.steps / (ms_total * 1e-3) / 1e9
    e2e_value = cells_step * args.steps / (ms_e2e * 1e-3) / 1e9
    peak, peak_src = peaks()
    per_launch_s = fused_ms * 1e-3 / max(fused_n, 1)
    achieved = E * n * n * ALG_BYTES_PER_CELL_UPDATE / per_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_cell_update")
        traffic = None if traffic is None else traffic * E * n * n

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{world * E} independent 700^2 WaveEnv (BASELINE configs[2] shard: {E}/GPU), triple-ring design "
                                   f"+ Gaussian source, {S} fused RK4 steps + energy trace per bench step",
                       "grid": [n, n], "envs_per_gpu": E, "rk4_steps_per_step": S, "parallelism": f"env-sharded x{world} (no collective)",
                       "l2": f"inputs larger than L2: {E * 12 * n * n * 4 / 1e9:.2f} GB of state per GPU streamed every RK4 step"},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT,
                    "h2d_bytes_per_step": int(world * (h_design.numel() * 4 + (S + 1) * 4)),
                    "d2h_bytes_per_step": int(world * h_energy.numel() * 4),
                    "note": "per env(action): design tables + tspan from host, energy signal to pinned host, env.wave frames stay on the device"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_fused_step<V> launch set (lean interior V=4 + PML strips V=1,2 + corners V=3): one RK4 step of the whole batch",
                         "traffic_note": "ncu dram bytes of one steady-state launch set (profiles/); below the 96 B/cell-update the metric counts because auxiliary fields that are constant in a window are not re-copied",
                         "launch_us": round(per_launch_s * 1e6, 1)},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args.cpu_steps)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def cpu_oracle_run(rk4_steps, threads=None):
    """The reference CPU path restated (oracle C port, OpenMP): one 700^2 environment of the workload."""
    from oracle import c_oracle as co
    from oracle import waves_oracle as wo
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    dim = wo.TwoDim.make(GRID_SIZE, N_GRID)
    dyn = wo.AcousticDynamics.make(dim, wo.WATER, 2.0, 20000.0)
    rng = np.random.default_rng(0)
    ds = wo.build_triple_ring_design_space()
    d0 = ds.sample(rng)
    d1 = ds(d0, wo.build_action_space(d0, 0.25).sample(rng))
    shape = wo.build_normal(wo.build_grid(dim), np.array([[-10.0, 0.0]]), np.array([0.3]), np.array([1.0]))
    ts = wo.build_tspan(np.float32(0.0), np.float32(1e-5), rk4_steps)
    dO = np.float32(wo.get_dx(dim) * wo.get_dy(dim))
    u0 = np.zeros((12, N_GRID, N_GRID), np.float32)
    co.lib()
    t0 = time.perf_counter()
    co.integrate(dyn, u0, ts, 1e-5, dO, d0, d1, ts[0], ts[-1], shape=shape, freq=1000.0)
    dt = time.perf_counter() - t0
    return N_GRID * N_GRID * rk4_steps / dt / 1e9, dt, co.num_threads()


def cpu_baseline(rk4_steps):
    v, dt, thr = cpu_oracle_run(rk4_steps)
    return {"value": round(v, 5), "unit": UNIT, "cores": thr, "kind": "port",
            "sample": f"1 of the workload's 700^2 environments, {rk4_steps} RK4 steps + energy trace ({dt:.1f} s), "
                      "oracle C restatement (OpenMP, -ffp-contract=off); Julia is not installed so the reference itself cannot run"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, all host cores) on bounded samples of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None
    sample_steps = max(1, args.ref_rk4_steps)
    for _ in range(args.warmup):
        cpu_oracle_run(1)
    t0 = time.perf_counter()
    vals = [cpu_oracle_run(sample_steps) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = sum(x[0] for x in vals) / len(vals)
    thr = vals[0][2]
    E = args.envs_per_gpu
    return {
        "impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{world * E} independent 700^2 WaveEnv (BASELINE configs[2] shard: {E}/GPU), triple-ring design "
                               f"+ Gaussian source; each reference step is a bounded sample: 1 environment x {sample_steps} RK4 steps",
                   "grid": [N_GRID, N_GRID]},
        "cpu_baseline": {"value": round(v, 5), "unit": UNIT, "cores": thr, "kind": "port",
                         "sample": f"1 environment x {sample_steps} RK4 steps per step (oracle C port, OpenMP x{thr}); "
                                   "the Julia reference cannot run in this image"},
        "e2e": {"value": round(v, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=128)
    ap.add_argument("--rk4-steps", type=int, default=100)
    ap.add_argument("--cpu-steps", type=int, default=100, help="RK4 steps of the cpu_baseline sample")
    ap.add_argument("--ref-rk4-steps", type=int, default=20, help="RK4 steps per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
