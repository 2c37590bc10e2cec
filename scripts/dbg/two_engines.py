"""Developer aid: upper bound of what running a batch as G independent groups (own streams, no step barrier between groups)
would gain: G engines of E/G environments each, driven by G host threads, against one engine of E environments.
   python scripts/dbg/two_engines.py 128 2"""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import waves_b200 as wb

E, G = int(sys.argv[1]), int(sys.argv[2])
steps, n = 100, 700
dim = wb.TwoDim(15.0, n)
ds = wb.build_triple_ring_design_space()
shape = wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0])
tspan = wb.build_tspan(0.0, 1e-5, steps)


def make(ne, seed):
    eng = wb.Engine(dim.x, dim.y, 1531.0, 1e-5, 2.0, 20000.0, n_env=ne)
    rng = np.random.default_rng(seed)
    eng.set_source(shape, 1000.0)
    for e in range(ne):
        d0 = ds.rand(rng)
        d1 = ds(d0, wb.build_action_space(d0, 0.25).rand(rng))
        eng.set_design(d0.table(), d1.table(), tspan[0], tspan[-1], env=e)
    for _ in range(2):
        eng.integrate(tspan, wb.MODE_FUSED, energy=True)
    return eng


def timed(engs, reps=5):
    bar = threading.Barrier(len(engs) + 1)
    def work(eng):
        bar.wait()
        for _ in range(reps):
            eng.integrate(tspan, wb.MODE_FUSED, energy=True)
        bar.wait()
    th = [threading.Thread(target=work, args=(e,)) for e in engs]
    for t in th: t.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for t in th: t.join()
    ne = sum(e.n_env for e in engs)
    return ne * n * n * steps * reps / dt / 1e9


one = make(E, 0)
print(f"1 engine x {E}: {timed([one]):.2f} Gcell-updates/s", flush=True)
one.close()
grp = [make(E // G, 10 + g) for g in range(G)]
print(f"{G} engines x {E // G}: {timed(grp):.2f} Gcell-updates/s", flush=True)
