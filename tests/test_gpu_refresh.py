"""The stencil refresh (openrbc_b200/csrc/rebuild.cuh: k_stencil_refresh, k_stencil_movers).  The reference searches its k-d tree for
the r < 6 / 8 / 9 centroid stencils at every rebuild (voronoi.h:105-117); the device re-classifies the neighbours it recorded within
9 + 1 at the last full search and searches in full only the cells whose centroid has jumped.  The stencils must be the same SETS in
the same order at every rebuild: compared here against a context that always searches, on hot fixtures whose cells do jump."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN

pytestmark = pytest.mark.gpu


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return {k[3:]: v for k, v in g.items() if k.startswith("in_")}


@pytest.mark.parametrize("name,dt", [("sphere_r12", 1e-2), ("vesicle_ico0", 1e-2), ("branches_vesicle_ico0", 2e-4)])
def test_refreshed_stencils_equal_searched_stencils(name, dt):
    from openrbc_b200 import Simulation
    st = load(name)
    a, b = Simulation(st, dt=dt, kBT=0.22, seed=77), Simulation(st, dt=dt, kBT=0.22, seed=77)
    b.set_option("stencil_refresh", 0)
    for sim in (a, b):
        sim.set_option("nl_reuse", 0)
    movers = 0
    for _ in range(11):                       # 22 steps: eleven rebuilds, none of them renumbers the cells (Morton sort every 24)
        a.run_langevin(2); b.run_langevin(2)
        ca, cb = a.dump("stencil_counts"), b.dump("stencil_counts")
        np.testing.assert_array_equal(ca, cb)
        sa, sb = a.dump("stencil"), b.dump("stencil")
        n9 = ca[:, 2].astype(np.int64)
        live = np.arange(sa.shape[1])[None, :] < n9[:, None]
        np.testing.assert_array_equal(np.where(live, sa, -1), np.where(live, sb, -1))
        for what in ("cell_start_l", "cell_start_p"):
            np.testing.assert_array_equal(a.dump(what), b.dump(what))
    c7 = int(a.dump("counters")[7])
    movers, redone = c7 & ((1 << 40) - 1), c7 >> 40
    assert int(b.dump("counters")[7]) == 0
    assert redone <= 1, (movers, redone)      # the refresh is the path taken
    print(f"{name}: {movers} cells searched in full over eleven rebuilds, {redone} refreshes redone")
    a.close(); b.close()
