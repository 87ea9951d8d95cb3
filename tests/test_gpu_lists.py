"""Hit lists with a skin (openrbc_b200/csrc/pair_queue.cuh): the force evaluation after a rebuild records, per particle, every candidate
closer than cut + skin; the evaluations up to the next rebuild walk those lists and re-test every entry with the reference's exact
guards.  The reference has no such lists (it searches the centroid stencils at every step, compute_pairwise_fused.h:238-320): the
lists must not change a single hit.  Checked here: trajectories with and without lists are equal (bit for bit where no atomics are
involved), the lists are really used, and a skin that is too thin for the step is caught by the displacement bound."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return {k[3:]: v for k, v in g.items() if k.startswith("in_")}


def run(st, steps, dt=1e-2, **opts):
    from openrbc_b200 import Simulation
    sim = Simulation(st, dt=dt, kBT=0.22, seed=4242)
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.run_langevin(steps)
    out = [sim.download(s, "xvno") for s in (0, 1)]
    stats = sim.dump("nl_stats")
    sim.close()
    return out, stats


@pytest.mark.parametrize("name", ["sphere_r12", "vesicle_ico0", "branches_vesicle_ico0"])
def test_lists_do_not_change_the_trajectory(name):
    st = load(name)
    exact = len(st["px"]) == 0                 # protein -> lipid reactions arrive by atomics: their order is not fixed
    steps = 10                                 # rebuild every 2nd step: five builds, five walks
    # the fixtures are freshly initialised membranes, far from equilibrium (forces of 1e2 .. 1e4): with the production time step
    # their fastest particles outrun the skin and the gate (rightly) orders searches; a short step keeps the walk legal
    dt = 1e-3 if name != "branches_vesicle_ico0" else 1e-5
    ref, s0 = run(st, steps, dt, nl_reuse=0)
    assert s0[0] == 0 and s0[1] == 0
    got, s1 = run(st, steps, dt, nl_reuse=1)
    assert s1[0] == steps // 2 and s1[1] == steps // 2 and s1[2] == 0, s1
    for s in (0, 1):
        for f in "xvno":
            if exact:
                np.testing.assert_array_equal(got[s][f], ref[s][f], err_msg=f)
            else:
                assert rel_err(got[s][f], ref[s][f]) < 1e-5, (s, f)
    # a skin thinner than twice the largest step (and no leave to thicken it): the gate never walks -- and does not bother to record either: every evaluation is a
    # plain search (the fourth counter), and nothing changes
    thin, s2 = run(st, steps, dt, nl_reuse=1, nl_skin=1e-7, nl_skin_max=1e-7)
    assert s2[0] <= 1 and s2[1] == 0 and s2[0] + s2[3] == steps, s2      # (the very first evaluation has seen no step yet)
    for s in (0, 1):
        for f in "xvno":
            if exact:
                np.testing.assert_array_equal(thin[s][f], ref[s][f], err_msg=f)
            else:
                assert rel_err(thin[s][f], ref[s][f]) < 1e-5, (s, f)


def test_lists_call_by_call_and_forces():
    """The call-by-call API: rebuild -> forces (build) -> integrate -> forces (walk); the walked forces equal a fresh search's."""
    from openrbc_b200 import Simulation
    st = load("vesicle_ico0")
    a, b = Simulation(st, dt=1e-3, kBT=0.0), Simulation(st, dt=1e-3, kBT=0.0)
    b.set_option("nl_reuse", 0)
    for sim in (a, b):
        sim.nstep = 24
        sim.rebuild(); sim.compute_pairwise_fused(); sim.compute_bonded(); sim.verlet_langevin()
        sim.compute_pairwise_fused()
    assert list(a.dump("nl_stats")[:2]) == [1, 1]
    for s in (0, 1):
        da, db = a.download(s, "ft"), b.download(s, "ft")
        assert rel_err(da["f"], db["f"]) < 1e-6 and rel_err(da["t"], db["t"]) < 1e-6
    # a position overwritten by the host invalidates the lists
    x = a.get(0, "x"); x[5] += 0.3
    a.set_field(0, "x", x); b.set_field(0, "x", x)
    for sim in (a, b):
        sim.clear_force(); sim.compute_pairwise_fused()
    assert list(a.dump("nl_stats")[:2]) == [2, 1]
    assert rel_err(a.get(0, "f"), b.get(0, "f")) < 1e-6
    a.close(); b.close()


def test_gate_thickens_the_skin_for_a_hot_system():
    """The fastest particles of the freshly initialised sphere outrun a skin of 0.002 even with a tenth of the production time step;
    allowed up to 0.3 the gate records thicker lists and walks them -- and the trajectory is still the one of the plain search."""
    st = load("sphere_r12")
    ref, _ = run(st, 10, 1e-3, nl_reuse=0)
    thin, s_thin = run(st, 10, 1e-3, nl_reuse=1, nl_skin=0.002, nl_skin_max=0.002)
    thick, s_thick = run(st, 10, 1e-3, nl_reuse=1, nl_skin=0.002, nl_skin_max=0.3)
    assert s_thick[1] > s_thin[1], (s_thin, s_thick)
    assert s_thick[2] == 0
    for got in (thin, thick):
        for s in (0, 1):
            for f in "xvno":
                np.testing.assert_array_equal(got[s][f], ref[s][f], err_msg=f)


@pytest.mark.parametrize("name", ["sphere_r12", "vesicle_ico0"])
def test_rows_that_overflow_fall_back_to_the_search(name):
    """List rows too short for the partners of a particle: the recording raises the overflow flag, the evaluations that follow search,
    the gate backs off -- and nothing changes in the trajectory."""
    st = load(name)
    exact = len(st["px"]) == 0
    ref, _ = run(st, 10, 1e-3, nl_reuse=0)
    got, stats = run(st, 10, 1e-3, nl_reuse=1, debug_nl_cap=4)
    assert stats[1] == 0 and stats[0] >= 1 and stats[0] + stats[3] == 10, stats     # recorded (in vain), never walked
    for s in (0, 1):
        for f in "xvno":
            if exact:
                np.testing.assert_array_equal(got[s][f], ref[s][f], err_msg=f)
            else:
                assert rel_err(got[s][f], ref[s][f]) < 1e-5, (s, f)
