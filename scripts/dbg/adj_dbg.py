import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import waves_b200 as wb
dim = wb.TwoDim(15.0, 700)
eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 2.0, 20000.0)
eng.set_source(wb.build_normal(dim, [[-10.0, 0.0]], [0.3], [1.0]), 1000.0)
rng = np.random.default_rng(0)
ds = wb.build_triple_ring_design_space()
d0 = ds.rand(rng)
ts = wb.build_tspan(0.0, 1e-5, 4)
eng.set_design(d0.table(), d0.table(), ts[0], ts[-1])
z0 = (rng.standard_normal((1, 12, 700, 700)) * 1e-3).astype(np.float32)
w1 = np.zeros((5, 3), np.float32); w1[-1, 2] = 1.0
mode = sys.argv[1] if len(sys.argv) > 1 else "adj"
eng.set_state(z0)
if mode == "int":
    print(eng.integrate(ts, wb.MODE_FUSED)[0][0, -1])
elif mode == "intne":
    print(eng.integrate(ts, wb.MODE_FUSED, energy=False))
elif mode == "step":
    for t in ts[:-1]:
        eng.step(float(t))
    print(eng.energy())
elif mode == "adjx":
    print(eng.adjoint(ts, w1, fwd_mode=wb.MODE_EXACT)[0])
else:
    print(eng.adjoint(ts, w1)[0])
