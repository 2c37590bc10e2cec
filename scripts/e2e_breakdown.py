"""Developer aid: where the end-to-end env(action) time of a large batch goes (wall vs device, frames kept or not).
   python scripts/e2e_breakdown.py [E]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb

E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n, S = 700, 100
dim = wb.TwoDim(15.0, n)
eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0, n_env=E, device=0)
stream = torch.cuda.ExternalStream(eng.stream(), device=0)
ds = wb.build_triple_ring_design_space()
rng = np.random.default_rng(0)
d0 = ds.rand(rng)
d1 = ds(d0, wb.build_action_space(d0, 0.25).rand(rng))
c0 = np.repeat(d0.table()[None], E, 0).astype(np.float32)
c1 = np.repeat(d1.table()[None], E, 0).astype(np.float32)
for e in range(E):
    eng.set_source(wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0]), 1000.0, env=e)
h_energy = torch.empty((E, S + 1, 3), dtype=torch.float32).pin_memory()
frames = torch.empty((E, 3, 12, n, n), dtype=torch.float32, device="cuda:0")
save = np.array([S - 20, S - 10, S], dtype=np.int32)
ts = wb.build_tspan(0.0, 1e-5, S)


def run(kind, reps=4):
    def once():
        t = {}
        a = time.perf_counter()
        eng.set_design_batch(c0, c1, ts[0], ts[-1])
        t["set_design"] = time.perf_counter() - a
        a = time.perf_counter()
        if kind == "frames":
            eng.integrate(ts, wb.MODE_FUSED, energy=h_energy.numpy(), save_steps=save, frames=frames)
        elif kind == "noenergy":
            eng.integrate(ts, wb.MODE_FUSED, energy=False)
        else:
            eng.integrate(ts, wb.MODE_FUSED, energy=h_energy.numpy())
        t["integrate"] = time.perf_counter() - a
        return t
    once(); once()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    ev0.record(stream)
    acc = {}
    for _ in range(reps):
        for k, v in once().items():
            acc[k] = acc.get(k, 0.0) + v
    ev1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) / reps * 1e3
    dev = ev0.elapsed_time(ev1) / reps
    print(f"{kind:9s} E={E}: wall {wall:8.2f} ms  events {dev:8.2f} ms   " + "  ".join(f"{k} {v / reps * 1e3:.2f} ms" for k, v in acc.items()), flush=True)


for kind in ("plain", "frames", "plain", "frames", "noenergy"):
    run(kind)
# the frame copies alone: 8 saved frames of an 8-step integration against none
ts8 = wb.build_tspan(0.0, 1e-5, 8)
fr8 = frames.view(-1)[: E * 3 * 12 * n * n].view(E, 3, 12, n, n)
for sv in ([], [2, 5, 8]):
    for _ in range(2):
        eng.integrate(ts8, wb.MODE_FUSED, energy=False, save_steps=sv, frames=fr8 if sv else None)
    torch.cuda.synchronize()
    a = time.perf_counter()
    for _ in range(5):
        eng.integrate(ts8, wb.MODE_FUSED, energy=False, save_steps=sv, frames=fr8 if sv else None)
    torch.cuda.synchronize()
    print(f"8 steps, {len(sv)} saved frames: {(time.perf_counter() - a) / 5 * 1e3:.2f} ms", flush=True)
eng.close()
