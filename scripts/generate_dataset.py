#!/usr/bin/env python
"""The dataset loop of the reference (scripts/data.jl:48-80: WaveEnv with the triple-ring design space and a
RandomPosGaussianSource at x = -10, integration_steps = 100, `actions` actions per episode, RandomDesignPolicy, one
`generate_episode!` + `save` per episode, 500 episodes under a 48-hour SLURM limit, scripts/data.sh:8) in its batched form:
the episodes of a rank step in lockstep on one handle (BatchWaveEnv), ranks take contiguous blocks of episodes with no
communication, every episode is written as episode<i>.npz or, with --bson, in the reference's BSON layout (src/data.jl:60-71).

  python scripts/generate_dataset.py --episodes 1024 --actions 20 --out /tmp/dataset [--bson] [--envs-per-batch 256]
  torchrun --nproc-per-node 8 scripts/generate_dataset.py --episodes 1024 ...
Prints one JSON line (rank 0): episodes, RK4 steps, wall seconds (max over ranks), episodes per hour, Gcell-updates/s.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=1024)
    ap.add_argument("--actions", type=int, default=20)
    ap.add_argument("--steps", type=int, default=100, help="integration_steps per action")
    ap.add_argument("--envs-per-batch", type=int, default=256, help="episodes that step in lockstep on one handle")
    ap.add_argument("--out", default=None, help="directory for the episode files (omit: episodes are generated and dropped)")
    ap.add_argument("--bson", action="store_true")
    ap.add_argument("--n", type=int, default=700)
    args = ap.parse_args()

    import torch
    import waves_b200 as wb
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    mine = wb.shard_envs(args.episodes, rank, world)
    dim = wb.TwoDim(15.0, args.n)
    ds = wb.build_triple_ring_design_space()
    if args.out and rank == 0:
        os.makedirs(args.out, exist_ok=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    written = 0
    for b0 in range(0, len(mine), args.envs_per_batch):
        ids = mine[b0:b0 + args.envs_per_batch]
        # episode i is seeded by its global index, whatever the sharding
        rngs = [np.random.default_rng(i) for i in ids]
        sources = [wb.RandomPosGaussianSource(dim, [-10.0, -10.0], [-10.0, 10.0], [0.3], [1.0], 1000.0, rng=r) for r in rngs]
        env = wb.BatchWaveEnv(dim, ds, sources, integration_steps=args.steps, actions=args.actions, device=local, rngs=rngs)
        policies = [wb.RandomDesignPolicy(sp, rng=r) for sp, r in zip(env.action_spaces(), rngs)]
        eps = wb.generate_episodes(policies, env)
        if args.out:
            for i, ep in zip(ids, eps):
                if args.bson:
                    ep.save_bson(os.path.join(args.out, f"episode{i + 1}.bson"), dim)
                else:
                    ep.save(os.path.join(args.out, f"episode{i + 1}.npz"))
                written += 1
        env.engine.close()
        del env
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    secs = float(dt.item())
    if rank == 0:
        cells = args.episodes * args.actions * args.steps * args.n * args.n
        print(json.dumps({"workload": f"{args.episodes} episodes x {args.actions} actions x {args.steps} RK4 steps, {args.n}^2 grid, "
                                      f"triple-ring design space, RandomPosGaussianSource, observations + energy signals",
                          "n_gpus": world, "episodes": args.episodes, "seconds": round(secs, 2),
                          "episodes_per_hour": round(args.episodes / secs * 3600.0, 1),
                          "Gcell_updates_per_s": round(cells / secs / 1e9, 2), "files": args.out is not None,
                          "format": "bson" if args.bson else "npz", "envs_per_batch": args.envs_per_batch,
                          "reference": "scripts/data.jl: one environment at a time, 500 episodes under a 48 h SLURM limit (scripts/data.sh:8)"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
