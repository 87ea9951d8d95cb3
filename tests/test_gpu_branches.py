"""CUDA path (through the C ABI) against the branch-coverage golden vectors (tests/golden/branches_*.npz, produced by the
reference itself with tests/golden/make_golden_branches.py).  The plain vesicle fixtures never reach these code paths:

  * protein-lipid Lennard-Jones            pairwise_kernel_fused.h:63-77, compute_pairwise_fused.h:174
  * protein-protein repulsion + LJ         pairwise_kernel_fused.h:79-97, compute_pairwise_fused.h:196-208 (type12 = type1 + 6 type2)
  * the reflecting wall                    integrate_nh.h:124-144 (bounce_back) and :200-209 (inside the fused initial kernel)
  * unfused verlet_nh_final / _update      integrate_nh.h:66-94,155-176
  * Langevin step with injected noise      integrate_langevin.h:99-149 with the reference's own xorshift stream
  * delete_lipid with survivors < N        cleanup.h:29-91

Every test first asserts that the branch it is about actually fires in the fixture (hit counts recomputed from the input state)."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN, branch_hits, rel_err

pytestmark = pytest.mark.gpu

F_TOL, X_TOL, T_TOL = 1e-4, 1e-6, 1e-5


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def sub(g, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in g.items() if k.startswith(prefix) and k[n:] not in ("lf", "lt", "pf", "pt")}


@pytest.mark.parametrize("pair_impl", [2, 1])
def test_pair_branches_fire_and_match(pair_impl):
    from openrbc_b200 import Simulation
    g = load("branches_vesicle_ico0")
    st = sub(g, "in_")
    hits = branch_hits(st, g, g["forcefield"])
    for k in ("ll", "pl_poly", "pl_lj", "pp_rep", "pp_lj"):
        assert hits[k] > 0, (k, hits)
    assert hits["pp_rep_type_pairs"] == 6 and hits["pp_lj_type_pairs"] == 4
    sim = Simulation(st, kBT=0.0)
    sim.set_option("pair_impl", pair_impl)
    sim.compute_pairwise_fused()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "ft")
        assert rel_err(d["f"], g[f"pair_{p}f"]) < F_TOL, (p, "f")
        assert rel_err(d["t"], g[f"pair_{p}t"]) < F_TOL, (p, "t")
    # the particles of the rare branches, one by one (the rms in rel_err must not hide them): a protein that takes part in a
    # protein-protein or LJ pair carries a force dominated by that pair
    ty = st["ptype"]
    d = sim.download(1, "f")["f"]
    ref = g["pair_pf"]
    big = np.linalg.norm(ref, axis=1) > 5.0
    assert big.sum() >= 20
    err = np.linalg.norm(d[big] - ref[big], axis=1) / np.linalg.norm(ref[big], axis=1)
    assert err.max() < 1e-4, (err.max(), ty[big][err.argmax()])
    for s, p in ((0, "l"), (1, "p")):
        sim.set_field(s, "f", g[f"pair_{p}f"]); sim.set_field(s, "t", g[f"pair_{p}t"])
    sim.compute_bonded()
    assert rel_err(sim.get(1, "f"), g["bonded_pf"]) < 2e-6
    sim.set_field(1, "f", g["bonded_pf"])
    sim.verlet_langevin()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xvno")
        for f in "xvno":
            assert rel_err(d[f], g[f"lang_{p}{f}"]) < X_TOL, (p, f)
    sim.close()
    # rebuild (no Morton step) from the reference's own post-step state: integer structures equal
    st2 = dict(st)
    for p in "lp":
        for f in "xvno":
            st2[p + f] = g[f"lang_{p}{f}"]
    sim = Simulation(st2, kBT=0.0)
    sim.nstep = 2
    sim.rebuild()
    np.testing.assert_array_equal(sim.dump("centroids"), g["rb_centroids"])
    for p in "lp":
        np.testing.assert_array_equal(sim.dump("aff_" + p), g[f"rb_aff_{p}"])
        np.testing.assert_array_equal(sim.dump("cell_start_" + p), g[f"rb_cs_{p}"])
        np.testing.assert_array_equal(sim.dump("cells_" + p), g[f"rb_cells_{p}"])
    d = sim.download(1, "x", ids=True)
    np.testing.assert_array_equal(d["type"], g["rb_ptype"]); np.testing.assert_array_equal(d["tag"], g["rb_ptag"])
    sim.close()


def test_unfused_nose_hoover_and_walls():
    from openrbc_b200 import Simulation
    g = load("branches_vesicle_ico0")
    assert int(g["hits_bounce_plain"]) > 20 and int(g["hits_bounce_fused"]) > 20
    st = sub(g, "nh_in_")
    sim = Simulation(st, kBT=0.22)
    sim.zeta = 0.04
    for s, p in ((0, "l"), (1, "p")):
        sim.set_field(s, "f", g[f"nh_in_{p}f"]); sim.set_field(s, "t", g[f"nh_in_{p}t"])
    sim.post_torque()
    for s, p in ((0, "l"), (1, "p")):
        assert rel_err(sim.get(s, "t"), g[f"pt_{p}t"]) < T_TOL
        sim.set_field(s, "t", g[f"pt_{p}t"])
    sim.nh_final()                                                 # verlet_nh_final
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "vo")
        for f in "vo":
            assert rel_err(d[f], g[f"nhfinal_{p}{f}"]) < X_TOL, (p, f)
            sim.set_field(s, f, g[f"nhfinal_{p}{f}"])
    sim.nh_update()                                                # verlet_nh_update: KE in fp64, zeta and Q on the host
    assert abs(sim.zeta - float(g["nhupd_zeta"])) <= 1e-6 * abs(float(g["nhupd_zeta"]))
    assert sim.Q.value == float(g["nhupd_Q"])
    # bounce_back alone: the box is drawn through the vesicle, all six faces fold particles back
    n_out = sum(int((np.abs(st[p + "x"]) > float(g["bb_box"])).any(axis=1).sum()) for p in "lp")
    assert n_out > 20
    sim.box = (-float(g["bb_box"]), float(g["bb_box"]))
    sim.bounce_back()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xv")
        np.testing.assert_array_equal(d["x"], g[f"bb_{p}x"])       # reflections are exact in fp32 / fp64 on both sides
        np.testing.assert_array_equal(d["v"], g[f"bb_{p}v"])
    # fused initial kernel, tighter box: drift, reflection, KE, zeta
    sim.zeta = float(g["nhupd_zeta"])
    sim.box = (-float(g["nhi_box"]), float(g["nhi_box"]))
    sim.nh_initial_fused()
    assert abs(sim.zeta - float(g["nhi_zeta"])) <= 1e-6 * abs(float(g["nhi_zeta"]))
    reflected = 0
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xvnoft")
        for f in "xvno":
            assert rel_err(d[f], g[f"nhi_{p}{f}"]) < X_TOL, (p, f)
        assert not d["f"].any() and not d["t"].any()
        # the reflected components themselves (lipids of the six caps; the proteins were all folded inside by the first box)
        out = np.abs(g[f"bb_{p}x"]) > float(g["nhi_box"]) + 0.01
        reflected += int(out.sum())
        if out.any():
            assert np.abs(d["x"][out] - g[f"nhi_{p}x"][out]).max() < 2e-5 and (np.abs(d["x"][out]) <= float(g["nhi_box"])).all()
            assert (np.sign(d["v"][out]) == np.sign(g[f"nhi_{p}v"][out])).all()
    assert reflected > 20
    sim.close()


def test_langevin_step_with_the_references_noise():
    """orbc_step_params.noise_*: the device integrator fed the very noise vectors the reference's MT19937 -> xorshift128 stream
    produced for this step (1 thread), so a thermal Langevin step can be compared with the reference value by value."""
    from openrbc_b200 import Simulation
    g = load("branches_vesicle_ico0")
    sim = Simulation(sub(g, "ln_in_"), kBT=0.22)
    for s, p in ((0, "l"), (1, "p")):
        sim.set_field(s, "f", g[f"ln_in_{p}f"]); sim.set_field(s, "t", g[f"ln_in_{p}t"])
    assert np.abs(g["ln_noise_l"]).max() > 0.9 and abs(g["ln_noise_l"].var() - 1 / 3) < 0.02
    sim.verlet_langevin(g["ln_noise_l"], g["ln_noise_p"])
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xvno")
        for f in "xvno":
            assert rel_err(d[f], g[f"ln_{p}{f}"]) < X_TOL, (p, f)
    # and the noise matters: without it the velocities differ by far more than the tolerance
    sim2 = Simulation(sub(g, "ln_in_"), kBT=0.0)
    for s, p in ((0, "l"), (1, "p")):
        sim2.set_field(s, "f", g[f"ln_in_{p}f"]); sim2.set_field(s, "t", g[f"ln_in_{p}t"])
    sim2.verlet_langevin()
    assert rel_err(sim2.get(0, "v"), g["ln_lv"]) > 1e-3
    sim.close(); sim2.close()


def test_delete_lipid_against_the_reference():
    """orbc_delete_lipid vs cleanup.h:29-91 run by the reference: survivors, their order, and the re-partition that follows."""
    from openrbc_b200 import Simulation
    g = load("branches_delete")
    st = sub(g, "del_in_")
    n0 = len(st["lx"])
    sim = Simulation(st, kBT=0.0)
    n1 = sim.delete_lipid(2.5)
    assert n1 == int(g["del_n"]) < n0 and sim.size(0) == n1
    d = sim.download(0, "xvnoft", affiliation=True)
    for f in "xvno":
        np.testing.assert_array_equal(d[f], g[f"del_l{f}"], err_msg=f)
    assert not d["f"].any() and not d["t"].any()
    np.testing.assert_array_equal(sim.dump("cell_start_l"), g["del_cs_l"])
    np.testing.assert_array_equal(d["affiliation"], np.repeat(np.arange(len(g["del_cs_l"]) - 1), np.diff(g["del_cs_l"])))
    assert sim.delete_lipid(2.5) == n1                             # nothing left at this tolerance: a no-op
    np.testing.assert_array_equal(sim.get(0, "x"), g["del_lx"])
    n2 = sim.delete_lipid(1.2)
    assert n2 == int(g["del2_n"]) < n1
    d = sim.download(0, "xvno")
    for f in "xvno":
        np.testing.assert_array_equal(d[f], g[f"del2_l{f}"], err_msg=f)
    np.testing.assert_array_equal(sim.dump("cell_start_l"), g["del2_cs_l"])
    # the shrunken system still runs: forces on the survivors are finite and the proteins kept their partition
    sim.compute_pairwise_fused(); sim.compute_bonded()
    assert np.isfinite(sim.get(0, "f")).all() and np.isfinite(sim.get(1, "f")).all()
    sim.close()


def test_upload_rejects_bad_ids():
    """The ids of an upload are checked on the device (k_check_ids, k_check_bonds): same refusals, same messages as a host sweep."""
    from openrbc_b200 import Simulation, engine
    g = dict(np.load(os.path.join(GOLDEN, "branches_vesicle_ico0.npz")))
    st = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    assert len(st["px"]) > 8 and len(st["bonds"]) > 4
    for field, slot, value, text in (("ptype", 5, 9, "protein 5 has type 9"), ("ptype", 3, -1, "protein 3 has type -1"), ("ptag", 2, -7, "negative protein tag")):
        bad = dict(st); bad[field] = np.array(st[field], np.int32); bad[field][slot] = value
        with pytest.raises(engine.OrbcError, match=text):
            Simulation(bad, kBT=0.0)
    for col, value, text in ((0, 4, "bond 2 has type 4"), (1, 10 ** 8, "bond 2 refers to a tag"), (2, -1, "bond 2 refers to a tag")):
        bad = dict(st); bad["bonds"] = np.array(st["bonds"], np.int32).reshape(-1, 3).copy(); bad["bonds"][2, col] = value
        with pytest.raises(engine.OrbcError, match=text):
            Simulation(bad, kBT=0.0)
    sim = Simulation(st, kBT=0.0)                # and the good state still loads
    sim.rebuild(); sim.compute_pairwise_fused(); sim.compute_bonded()
    assert np.isfinite(sim.get(1, "f")).all()
    sim.close()
