"""Developer aid: one fused step on a small grid; WAVES_DEBUG_SKIP=1|2 skips the interior|general kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import waves_b200 as wb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dim = wb.TwoDim(3.0, n)
eng = wb.Engine(dim.x, dim.y, 1531.0, 1e-5, 0.6, 20000.0)
u0 = (np.random.default_rng(0).standard_normal((1, 12, n, n)) * 1e-3).astype(np.float32)
eng.set_state(u0)
try:
    eng.step(0.0, wb.MODE_FUSED)
    out = eng.get_state(0)
    print("OK skip=%s" % os.environ.get("WAVES_DEBUG_SKIP"), np.isfinite(out).all(), np.abs(out - u0[0]).max())
except Exception as e:
    print("FAIL skip=%s" % os.environ.get("WAVES_DEBUG_SKIP"), str(e)[:150])
