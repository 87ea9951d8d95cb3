"""CUDA path against the second set of golden vectors (tests/golden/ext_*.npz, produced by the reference itself with
tests/golden/make_golden_ext.py): trajectory frames, the energy-minimisation loop and the volume constraint."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

FILES = ["vesicle_ico0", "sphere_r12"]


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, "ext_" + name + ".npz")))
    return g, {k[3:]: v for k, v in g.items() if k.startswith("in_")}


@pytest.mark.parametrize("name", FILES)
def test_save_frame_bytes(name):
    """The frame assembled on the device == the file the reference's save_frame writes, byte for byte (trajectory.h:61-105);
    the asynchronous double-buffered path returns the same bytes."""
    from openrbc_b200 import Simulation
    g, st = load(name)
    sim = Simulation(st, kBT=0.0)
    sim.nstep = int(g["frame_nstep"])
    for s, p in ((0, "l"), (1, "p")):
        sim.set_field(s, "f", g[f"frc_{p}f"])
    base = int(g["tag_base"])
    for df in (7, 31, 1):
        want = g[f"frame_{df}"]
        assert sim.frame_bytes(df) == len(want)
        np.testing.assert_array_equal(sim.save_frame(df, base), want)
    sim.save_frame_begin(7, base); sim.save_frame_begin(31, base)
    with pytest.raises(Exception):
        sim.save_frame_begin(1, base)                  # two in flight already
    np.testing.assert_array_equal(sim.save_frame_end(), g["frame_7"])
    sim.save_frame_begin(1, base)
    np.testing.assert_array_equal(sim.save_frame_end(), g["frame_31"])
    np.testing.assert_array_equal(sim.save_frame_end(), g["frame_1"])
    sim.close()


@pytest.mark.parametrize("name", FILES)
def test_constrain_volume(name):
    from openrbc_b200 import Simulation
    g, st = load(name)
    sim = Simulation(st, kBT=0.0)
    for k in (1, 2):
        sim.clear_force()
        vol = sim.constrain_volume(3.15, 0.05)
        assert np.isfinite(vol)
        for s, p in ((0, "l"), (1, "p")):
            assert rel_err(sim.get(s, "f"), g[f"cv{k}_{p}f"]) < 1e-5, (k, p)
    sim.close()


@pytest.mark.parametrize("name", FILES)
@pytest.mark.parametrize("fused", [False, True])
def test_minimisation_loop(name, fused):
    """Two iterations of openrbc.cpp:88-133: call by call (clear_force, forces, post_torque, mover, bounce_back) and through the
    whole-loop entry point with the fused kernel."""
    from openrbc_b200 import Simulation
    g, st = load(name)
    sim = Simulation(st, kBT=0.0)
    if fused:
        sim.run_minimize(2)
    else:
        for _ in range(2):
            sim.nstep = 0
            sim.rebuild()
            sim.clear_force(); sim.compute_pairwise_fused(); sim.compute_bonded()
            sim.post_torque(); sim.opt_move(); sim.bounce_back()
    np.testing.assert_array_equal(sim.dump("cell_start_l"), g["opt_cs_l"])
    np.testing.assert_array_equal(sim.dump("cell_start_p"), g["opt_cs_p"])
    assert rel_err(sim.dump("centroids"), g["opt_centroids"]) < 1e-6
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xnft")
        assert rel_err(d["x"], g[f"opt_{p}x"]) < 1e-6 and rel_err(d["n"], g[f"opt_{p}n"]) < 1e-6
        assert rel_err(d["f"], g[f"opt_{p}f"]) < 1e-4
        assert rel_err(d["t"], g[f"opt_{p}t"]) < 1e-4      # n x t of the last iteration
    sim.close()


def test_volume_constraint_inside_run_langevin():
    """orbc_set_volume_constraint: run_langevin with the constraint == the call-by-call loop with constrain_volume at openrbc.cpp:229."""
    from openrbc_b200 import Simulation
    g, st = load("vesicle_ico0")
    a, b = Simulation(st, kBT=0.0), Simulation(st, kBT=0.0)
    a.set_volume_constraint(True, 3.15, 0.05)
    a.run_langevin(4)
    for _ in range(4):
        if b.nstep % b.freq_voronoi == 0:
            b.rebuild()
        b.compute_pairwise_fused(); b.compute_bonded(); b.constrain_volume(3.15, 0.05)
        b.verlet_langevin(); b.nstep += 1
    c = Simulation(st, kBT=0.0); c.run_langevin(4)
    for s in (0, 1):
        da, db, dc = a.download(s, "xv"), b.download(s, "xv"), c.download(s, "xv")
        for f in "xv":
            assert rel_err(da[f], db[f]) < 1e-6
        assert rel_err(da["v"], dc["v"]) > 1e-6            # and the constraint does act
    for sim in (a, b, c):
        sim.close()


@pytest.mark.parametrize("name", ["sphere_r12", "sphere_r16"])
def test_voronoi_init(name):
    """orbc_voronoi_init == VoronoiDiagram::init of the reference (voronoi.h:54-75): centroids, cell_start and the storage order
    of the container, bit for bit (the summation order of every centroid update is the reference's)."""
    from openrbc_b200 import Simulation
    g = dict(np.load(os.path.join(GOLDEN, "init_" + name + ".npz")))
    n = len(g["x0"])
    z = np.zeros((n, 3), np.float32)
    e = np.zeros((0, 3), np.float32)
    st = dict(lx=g["x0"], lv=z, ln=g["n0"], lo=z, px=e, pv=e, pn=e, po=e, ptype=np.zeros(0, np.int32), ptag=np.zeros(0, np.int32), bonds=np.zeros((0, 3), np.int32))
    sim = Simulation(st, kBT=0.0)
    sim.voronoi_init(int(g["n_cells"]), int(g["n_iter"]))
    np.testing.assert_array_equal(sim.dump("centroids"), g["centroids"])
    np.testing.assert_array_equal(sim.dump("cell_start_l"), g["cs_l"])
    d = sim.download(0, "xn", affiliation=True)
    np.testing.assert_array_equal(d["x"], g["x"])
    np.testing.assert_array_equal(d["n"], g["n"])
    np.testing.assert_array_equal(d["affiliation"], np.repeat(np.arange(len(g["cs_l"]) - 1), np.diff(g["cs_l"])))
    # the state is ready for the loop: forces can be computed straight away
    sim.compute_pairwise_fused()
    assert np.isfinite(sim.get(0, "f")).all()
    with pytest.raises(Exception):
        sim.voronoi_init(int(g["n_cells"]), 5)             # n_iterate - 1 a power of two: the reference's double reorder
    sim.close()
