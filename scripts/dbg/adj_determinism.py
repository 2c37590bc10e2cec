"""Developer aid: is the fused reverse pass deterministic?  Repeats one small gradient and counts runs that differ from the first."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import waves_b200 as wb
from test_gpu_adjoint import setup
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for n, pw in ((200, 0.25), (131, 0.2)):
    for mode in (wb.ADJ_EXACT, wb.ADJ_COMPAT):
        p, eng, ts, z0, w, aN = setup(n=n, steps=7, pml_width=pw)
        for march in (True, False):
            outs = []
            for r in range(reps):
                eng.set_state(z0[None])
                _, g, _ = eng.adjoint(ts, w, aN[None], adj_mode=mode, want_dc=False, march=march)
                outs.append(g.copy())
            bad = [r for r in range(1, reps) if not np.array_equal(outs[r], outs[0])]
            worst = max([float(np.abs(outs[r] - outs[0]).max()) for r in bad], default=0.0)
            fields = sorted({int(f) for r in bad for f in range(12) if not np.array_equal(outs[r][0, f], outs[0][0, f])})
            print(f"n={n} mode={mode} march={march}: {len(bad)} of {reps - 1} repeats differ, max abs diff {worst:.3e}, fields {fields}", flush=True)
        eng.close()
