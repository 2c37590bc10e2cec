"""Reverse pass without dL/dc (the gradient the reference itself produces: its cylinder mask has no derivative): the forward
stage states are not recomputed (the dynamics are linear in the state, J^T does not depend on them), so a reverse step is four
launches instead of seven.  Written after the round's GPU budget was spent, hence in a file that sorts last: the result must
be bitwise the dL/dz0 of the full sweep."""
import numpy as np
import pytest

import waves_b200 as wb
from test_gpu_adjoint import rel, setup, TOL
from oracle import adjoint_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("adj_mode", [wb.ADJ_EXACT, wb.ADJ_COMPAT])
def test_z0_gradient_without_dc_is_bitwise_the_full_sweep(adj_mode):
    p, eng, ts, z0, w, aN = setup(steps=6)
    eng.set_state(z0[None])
    n0 = eng.launch_count()
    _, dz_full, dc = eng.adjoint(ts, w, aN[None], fwd_mode=wb.MODE_EXACT, adj_mode=adj_mode)
    n1 = eng.launch_count()
    eng.set_state(z0[None])
    _, dz_only, none = eng.adjoint(ts, w, aN[None], fwd_mode=wb.MODE_EXACT, adj_mode=adj_mode, want_dc=False)
    n2 = eng.launch_count()
    assert none is None and dc is not None
    assert np.array_equal(dz_only, dz_full)
    sweeps = 6 if adj_mode == wb.ADJ_EXACT else 7
    assert (n1 - n0) - (n2 - n1) == 3 * sweeps          # three forward-stage launches per reverse step are gone
    if adj_mode == wb.ADJ_EXACT:
        _, gz, _, _ = ao.autograd_truth(p, z0, w, aN)
        assert rel(dz_only[0], gz) < TOL
    eng.close()
