#!/usr/bin/env python
"""bench.py -- fused 2D PML RK4 wave step: Gcell-updates/s on 1/2/4/8 B200 vs the HBM roofline.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): a batch of 1024 independent 700^2 WaveEnv
environments (triple-ring design + Gaussian source) sharded over the GPUs with no communication: 1024 on one GPU, 128 each on
eight ("scaling": "strong").  `--envs-per-gpu E` fixes the per-GPU batch instead ("weak").  One bench "step" = one
`env(action)` for every environment of the batch: 100 fused RK4 steps + the 101x3 energy trace (src/env.jl:91-121).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--total-envs T | --envs-per-gpu E] [--rk4-steps S]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)
  python bench.py --impl reference ...                      (reference CPU path: the oracle port, all host cores)

Sub-records after the headline (each bounded to a few seconds; --no-extras skips them):
  "weak"        the 128-environments-per-GPU shard at N = 1 (the round-1 headline, for continuity)
  "single_env"  BASELINE configs[1]: one 700^2 WaveEnv, us per RK4 step of env(action) (N = 1)
  "adjoint"     BASELINE configs[4]: gradient of the scattered energy through 500 steps, forward + reverse, one environment and a
                batch of 32 x 100 steps (N = 1)
  "slab"        BASELINE configs[3]: one 16384^2 grid slab-decomposed over the N GPUs, NVLink peer-memory halo (N > 1)
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


def measure_batch(wb, torch, dist, rank, world, local, E, env0, total_envs, S, steps, warmup, sample_clocks=True):
    """The headline measurement on this rank's E environments (global indices env0 ..): value, e2e, live kernel timing."""
    n = N_GRID
    dim = wb.TwoDim(GRID_SIZE, n)
    make_env_inputs.ds = wb.build_triple_ring_design_space()
    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=E, device=local)
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)

    # synthetic inputs: per-env design pair + source position; state starts at rest like reset!(env)
    designs = []
    for e in range(E):
        d0, d1, mu_y = make_env_inputs(env0 + e)
        designs.append((d0.table(), d1.table()))
        eng.set_source(wb.build_normal(dim, [[-10.0, mu_y]], [0.3], [1.0]), 1000.0, env=e)
    # pinned host staging for the per-step inputs (design tables) and outputs (energy signal)
    ncyl = designs[0][0].shape[0]
    h_design = torch.empty((E, 2, ncyl, 4), dtype=torch.float32).pin_memory()
    for e in range(E):
        h_design[e, 0] = torch.from_numpy(designs[e][0])
        h_design[e, 1] = torch.from_numpy(designs[e][1])
    h_energy = torch.empty((E, S + 1, 3), dtype=torch.float32).pin_memory()
    # env.wave (frames 80, 90, 100; src/env.jl:116) stays on the device; with little free memory only the last frame is kept
    free_b = torch.cuda.mem_get_info(local)[0]
    nkeep = 3 if (S >= 20 and free_b > E * 3 * 12 * n * n * 4 + 24e9) else 1
    d_frames = torch.empty((E, nkeep, 12, n, n), dtype=torch.float32, device=f"cuda:{local}")
    save = np.array([S - 20, S - 10, S][3 - nkeep:], dtype=np.int32)
    hd = h_design.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    time_step = [0]

    def one_step(e2e):
        """One env(action) for the whole batch: the DesignInterpolator of every environment is rebound to the new tspan (the
        design moves during every timed step), 100 fused RK4 steps, energy signal to pinned host memory.  e2e additionally
        keeps env.wave (the saved frames) on the device like the reference's gpu(env)."""
        ts = wb.build_tspan(np.float32(np.float32(time_step[0]) * np.float32(1e-5)), 1e-5, S)
        eng.set_design_batch(hd[:, 0], hd[:, 1], ts[0], ts[-1])
        if e2e:
            eng.integrate(ts, wb.MODE_FUSED, energy=h_energy.numpy(), save_steps=save, frames=d_frames)
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
            ms = wall * 1e3  # end to end is a host-visible latency: wall clock between the barriers
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.launch_count() - l0

    for _ in range(warmup):
        one_step(False)
    sampler = ClockSampler(local)
    if rank == 0 and sample_clocks:
        sampler.start()
    ms_total, launches = timed(steps, False)
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    replays0 = eng.graph_replays()
    one_step(True)
    sampler2 = ClockSampler(local)
    if rank == 0 and sample_clocks:
        sampler2.start()
    ms_e2e, _ = timed(steps, True)
    clocks_e2e = sampler2.stop() if (rank == 0 and sample_clocks) else None
    graph_replays = eng.graph_replays() - replays0

    # dominant-kernel duration, measured live with CUDA events on the launching stream (separate pass, kernels launched
    # directly: the profiling events sit between the launch sets)
    eng.profile(True)
    one_step(False)
    fused_ms, fused_n = eng.profile_read()
    eng.profile(False)
    eng.close()
    del d_frames
    torch.cuda.empty_cache()

    cells_step = total_envs * n * n * S
    per_launch_s = fused_ms * 1e-3 / max(fused_n, 1)
    return {"value": cells_step * steps / (ms_total * 1e-3) / 1e9, "e2e": cells_step * steps / (ms_e2e * 1e-3) / 1e9,
            "ms_per_step": ms_total / steps, "launches": int(launches), "per_launch_s": per_launch_s, "clocks": clocks,
            "h2d": int(world * (h_design.numel() * 4 + (S + 1) * 4)), "d2h": int(world * h_energy.numel() * 4),
            "frames_kept": nkeep, "graph_replays_per_e2e_step": graph_replays / (steps + 1), "clocks_e2e": clocks_e2e}


def run_ours(args):
    import torch
    import waves_b200 as wb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if wb._lib.lib().waves_build_flags() != 0:
        raise SystemExit("bench.py refuses a developer build of libwaves_b200.so (WAVES_DEV: honours WAVES_DEBUG_* switches)")
    debug_env = sorted(k for k in os.environ if k.startswith("WAVES_DEBUG"))   # listed for the record: the release library reads none
    strong = args.envs_per_gpu is None
    if strong:
        my_envs = wb.shard_envs(args.total_envs, rank, world)
        E, env0, total_envs = len(my_envs), my_envs[0], args.total_envs
    else:
        E, env0, total_envs = args.envs_per_gpu, rank * args.envs_per_gpu, world * args.envs_per_gpu
    S, n = args.rk4_steps, N_GRID
    m = measure_batch(wb, torch, dist, rank, world, local, E, env0, total_envs, S, args.steps, args.warmup)

    peak, peak_src = peaks()
    achieved = E * n * n * ALG_BYTES_PER_CELL_UPDATE / m["per_launch_s"] / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("dram_bytes_per_cell_update")
        traffic = None if traffic is None else traffic * E * n * n
        traffic_src = "ncu, committed (profiles/ncu_traffic.json: " + tj.get("source", "") + ")"
    out = None
    if rank == 0:
        real = None if traffic is None else traffic / m["per_launch_s"] / 1e9
        out = {
            "metric": METRIC, "value": round(m["value"], 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(m["ms_per_step"], 3), "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{total_envs} independent 700^2 WaveEnv (BASELINE configs[2]" + (": the whole batch" if strong else " shard")
                                   + f", {E}/GPU), triple-ring design + Gaussian source, designs rebound (moving) every bench step, "
                                   f"{S} fused RK4 steps + energy trace per bench step",
                       "grid": [n, n], "total_envs": total_envs, "envs_per_gpu": E, "rk4_steps_per_step": S,
                       "parallelism": f"env-sharded x{world} (no collective)",
                       "l2": f"inputs larger than L2: {E * 12 * n * n * 4 / 1e9:.2f} GB of state per GPU streamed every RK4 step"},
            "e2e": {"value": round(m["e2e"], 3), "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "note": f"per env(action): design tables + tspan from host, energy signal to pinned host, env.wave ({m['frames_kept']} "
                            f"frame(s) x 12 fields) kept on the device; {m['graph_replays_per_e2e_step']:.2f} CUDA graph launches per step; "
                            "measured after the value window in the same process (the GPU is at its sustained, power-capped clock by then: "
                            "clocks_sm_mhz below against clocks.sm_mhz)",
                    "clocks_sm_mhz": None if not m.get("clocks_e2e") else m["clocks_e2e"]["sm_mhz"]},
            "gpu_launches": m["launches"],
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "achieved_real": None if real is None else round(real, 1), "frac_real": None if real is None else round(real / peak, 4),
                         "kernel": "k_fused_step<V> launch set (lean interior V=4 + PML strips V=1,2 + corners V=3): one RK4 step of the whole batch",
                         "note": "achieved/frac: the contract's 96 B per cell-update; achieved_real/frac_real: the DRAM bytes ncu measured for one "
                                 "steady-state launch set (fewer: auxiliary fields that are constant in a window are not re-copied) over the same "
                                 "live CUDA-event duration",
                         "launch_us": round(m["per_launch_s"] * 1e6, 1)},
            "clocks": m["clocks"], "debug_env": debug_env,
        }
    # ---- sub-records (bounded; every rank takes part where a process group is involved) ----
    if not args.no_extras:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        if world == 1:
            extras = {}
            if strong and E != 128:
                w = measure_batch(wb, torch, dist, rank, world, local, 128, 0, 128, S, 3, 3, sample_clocks=False)
                extras["weak"] = {"envs_per_gpu": 128, "value": round(w["value"], 3), "e2e": round(w["e2e"], 3), "unit": UNIT,
                                  "frac": round(128 * n * n * ALG_BYTES_PER_CELL_UPDATE / w["per_launch_s"] / 1e9 / peak, 4),
                                  "note": "the 8-GPU shard of the batch on one GPU (round-1 headline configuration)"}
            try:
                import bench_single_env
                se = bench_single_env.measure(actions=10, device=local)
                extras["single_env"] = {"workload": "BASELINE configs[1]: one 700^2 WaveEnv, env(action) = 100 RK4 steps + energy trace",
                                        "single_env_us_per_step": se["default"]["us_per_rk4_step"], **se}
            except Exception as ex:  # a sub-record must never take the headline down
                extras["single_env"] = {"error": repr(ex)[:300]}
            try:
                import bench_adjoint
                extras["adjoint"] = bench_adjoint.measure(steps=500, E=1, device=local)
                b32 = bench_adjoint.measure(steps=100, E=32, device=local, with_dc=False, check_exact=False)
                extras["adjoint"]["batch_32_envs_100_steps"] = {k: b32[k] for k in ("forward_only_seconds", "seconds", "value", "reverse_Gcell_per_s")}
            except Exception as ex:
                extras.setdefault("adjoint", {})["error"] = repr(ex)[:300]
            out.update(extras)
            out["single_env_us_per_step"] = extras.get("single_env", {}).get("single_env_us_per_step")
        else:
            try:
                import bench_slab
                rec = bench_slab.measure(16384, 20, True, rank, world, local)
            except Exception as ex:
                rec = {"error": repr(ex)[:300]}
            if rank == 0:
                out["slab"] = rec
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args.cpu_steps)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def cpu_oracle_run(rk4_steps, threads=None):
    """The reference CPU path restated (oracle C port, OpenMP): one 700^2 environment of the workload."""
    from oracle import c_oracle as co
    from oracle import waves_oracle as wo
    # every host core, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    co.set_num_threads(threads or os.cpu_count() or 1)
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
    t0 = time.perf_counter()
    co.integrate(dyn, u0, ts, 1e-5, dO, d0, d1, ts[0], ts[-1], shape=shape, freq=1000.0)
    dt = time.perf_counter() - t0
    return N_GRID * N_GRID * rk4_steps / dt / 1e9, dt, co.num_threads()


def cpu_ref_faithful_run(rk4_steps=3):
    """The same environment through the structurally faithful statement of the reference's CPU path (oracle/sparse_oracle.py:
    the derivative as a SciPy CSC matrix applied as grad * u / (grad * u')', every operation its own float32 broadcast, the
    speed mask rebuilt at every stage time like `C = t -> speed(interp(t), grid, c0)`): what Julia's CPU arrays do, without
    Julia's compiled broadcast fusion.  A few steps suffice: the cost per step is constant."""
    from oracle import sparse_oracle as so
    from oracle import waves_oracle as wo
    F32 = np.float32
    dyn = so.Dynamics(GRID_SIZE, N_GRID, wo.WATER, 2.0, 20000.0)
    rng = np.random.default_rng(0)
    ds = wo.build_triple_ring_design_space()
    d0 = ds.sample(rng)
    d1 = ds(d0, wo.build_action_space(d0, 0.25).sample(rng))
    def tab(d):   # stacked cylinders of the design (speed(::Cloak), src/designs.jl:228) as an (ncyl, 4) table {x, y, r, c}
        c = d.all_cylinders() if hasattr(d, "all_cylinders") else d
        return np.concatenate([c.pos, c.r[:, None], c.c[:, None]], 1).astype(F32)
    t0c, t1c = tab(d0), tab(d1)
    shape = so.build_normal(dyn.grid, np.array([[-10.0, 0.0]]), [0.3], [1.0])
    ts = so.jl_range(F32(0.0), F32(rk4_steps * 1e-5), rk4_steps + 1)

    def C(t):
        d = so.design_at(t, t0c, t1c, ts[0], ts[-1])
        return so.speed(d[:, :2], d[:, 2], d[:, 3], dyn.grid, dyn.c0)

    t0 = time.perf_counter()
    sol = so.integrate(dyn, np.zeros((N_GRID, N_GRID, 12), F32), ts, (C, so.source(shape, 1000.0)), 1e-5)
    so.energies(sol, F32(np.mean(np.diff(dyn.x))), F32(np.mean(np.diff(dyn.y))))
    dt = time.perf_counter() - t0
    return N_GRID * N_GRID * rk4_steps / dt / 1e9, dt


def cpu_baseline(rk4_steps):
    v, dt, thr = cpu_oracle_run(rk4_steps)
    try:
        vf, dtf = cpu_ref_faithful_run(3)
        faithful = {"value": round(vf, 6), "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"the same environment, 3 RK4 steps + energies ({dtf:.1f} s): SciPy CSC matrix products and unfused float32 "
                              "broadcasts, statement by statement like the reference's CPU arrays (oracle/sparse_oracle.py)"}
    except Exception as ex:   # the faithful arm is an extra; it must never take the baseline down
        faithful = {"error": repr(ex)[:200]}
    return {"ref_faithful": faithful, "value": round(v, 5), "unit": UNIT, "cores": thr, "kind": "port",
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
    strong = args.envs_per_gpu is None
    total = args.total_envs if strong else world * args.envs_per_gpu
    return {
        "impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 1), "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{total} independent 700^2 WaveEnv (BASELINE configs[2]), triple-ring design + Gaussian source; "
                               f"each reference step is a bounded sample: 1 environment x {sample_steps} RK4 steps on all host cores",
                   "grid": [N_GRID, N_GRID], "total_envs": total},
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
    ap.add_argument("--total-envs", type=int, default=1024, help="environments of the whole job, sharded over the GPUs (strong scaling)")
    ap.add_argument("--envs-per-gpu", type=int, default=None, help="fix the per-GPU batch instead (weak scaling)")
    ap.add_argument("--no-extras", action="store_true", help="skip the weak / single_env / adjoint / slab sub-records")
    ap.add_argument("--rk4-steps", type=int, default=100)
    ap.add_argument("--cpu-steps", type=int, default=100, help="RK4 steps of the cpu_baseline sample")
    ap.add_argument("--ref-rk4-steps", type=int, default=20, help="RK4 steps per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: anything else that writes to file descriptor 1 (the NCCL version banner
    # of the first communicator, library chatter) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)


if __name__ == "__main__":
    main()
