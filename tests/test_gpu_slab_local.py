"""Slab decomposition (BASELINE configs[3]) checked on ONE GPU: the slabs of a grid live in separate handles on
the same device and trade ghost rows through waves_halo_pack / waves_halo_unpack (waves_b200.LocalSlabGroup).  The result must
equal the single-handle run of the same grid bit for bit -- same bar as tests/test_gpu_multi.py, which needs >= 2 GPUs."""
import numpy as np
import pytest

import waves_b200 as wb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,world,steps,gs", [(512, 2, 10, 11.0), (512, 4, 8, 11.0), (257, 3, 9, 5.5)])   # 257: pitch != nx, uneven slabs
@pytest.mark.parametrize("mode", ["fused", "exact"])
def test_local_slabs_match_single_handle(n, world, steps, gs, mode):
    m = wb.MODE_FUSED if mode == "fused" else wb.MODE_EXACT
    dim = wb.TwoDim(gs, n)
    rng = np.random.default_rng(n + world)
    u0 = (rng.standard_normal((12, n, n)) * 1e-3).astype(np.float32)
    shape = wb.build_normal(dim, [[-0.3 * gs, 0.04 * gs]], [0.03 * gs], [1.0])
    ts = wb.build_tspan(0.0, 1e-5, steps)

    eng = wb.Engine(dim.x, dim.y, wb.WATER, 1e-5, 0.2 * gs, 20000.0, device=0)
    eng.set_state(u0[None])
    eng.set_source(shape, 1000.0)
    ren, _ = eng.integrate(ts, m)
    ref = eng.get_state(0)
    eng.close()

    grp = wb.LocalSlabGroup(dim.x, dim.y, wb.WATER, 1e-5, 0.2 * gs, 20000.0, devices=[0] * world)
    grp.set_source_global(shape, 1000.0)
    grp.set_state_global(u0)
    en = grp.integrate(ts, m)
    got = grp.gather_state()
    # a second run on the same handles from a fresh state (the constant-field shortcut must be invalidated on every slab)
    grp.set_state_global(u0)
    en2 = grp.integrate(ts, m)
    got2 = grp.gather_state()
    grp.close()
    assert np.array_equal(got, ref), f"slab run differs from the single handle: max |d| = {np.abs(got - ref).max()}"
    assert np.array_equal(got2, ref)
    assert np.abs(en - ren[0]).max() / ren[0].max() < 1e-6 and np.array_equal(en, en2)
