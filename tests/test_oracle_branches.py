"""The C restatement (oracle/orbc_oracle.c) against the branch-coverage golden vectors (tests/golden/branches_*.npz, produced by
the reference itself with tests/golden/make_golden_branches.py): Lennard-Jones and protein-protein branches of the pair driver,
the unfused Nose-Hoover kernels, the reflecting wall, a Langevin step with the reference's own noise, and delete_lipid with
survivors < N.  Runs anywhere: needs neither the reference tree nor a GPU."""
import os

import numpy as np

from oracle import port
from tests.common import GOLDEN, branch_hits, rel_err


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def sub(g, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in g.items() if k.startswith(prefix) and k[n:] not in ("lf", "lt", "pf", "pt")}


def test_every_pair_branch_fires_and_matches():
    g = load("branches_vesicle_ico0")
    st = sub(g, "in_")
    hits = branch_hits(st, g, g["forcefield"])
    for k in ("ll", "pl_poly", "pl_lj", "pp_rep", "pp_lj"):
        assert hits[k] > 0 and hits[k] == int(g["hits_" + k]), (k, hits)
    assert hits["pp_rep_type_pairs"] == 6 and hits["pp_lj_type_pairs"] == 4     # all of {1,2,3}^2; (1,4) (1,5) (2,4) (2,5)
    w = port.World(st, kBT=0.0)
    w.compute_pairwise_fused()
    for p in "lp":
        for f in "ft":
            assert rel_err(getattr(w, p + f), g[f"pair_{p}{f}"]) < 2e-5, (p, f)
    w.lf, w.lt, w.pf, w.pt = (g[k].copy() for k in ("pair_lf", "pair_lt", "pair_pf", "pair_pt"))
    w.compute_bonded()
    np.testing.assert_array_equal(w.pf, g["bonded_pf"])
    w.verlet_langevin()
    for p in "lp":
        for f in "xvno":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"lang_{p}{f}"], err_msg=p + f)
    w.nstep = 2
    w.rebuild()
    np.testing.assert_array_equal(w.centroids, g["rb_centroids"])
    for p in "lp":
        np.testing.assert_array_equal(getattr(w, "aff_" + p), g[f"rb_aff_{p}"])
        np.testing.assert_array_equal(getattr(w, "cs_" + p), g[f"rb_cs_{p}"])
        np.testing.assert_array_equal(getattr(w, "cells_" + p), g[f"rb_cells_{p}"])


def test_unfused_nose_hoover_wall_and_injected_noise():
    g = load("branches_vesicle_ico0")
    w = port.World(sub(g, "nh_in_"), kBT=0.22)
    w.zeta = 0.04
    w.lf, w.lt, w.pf, w.pt = (g[k].copy() for k in ("nh_in_lf", "nh_in_lt", "nh_in_pf", "nh_in_pt"))
    w.post_torque()
    for p in "lp":
        np.testing.assert_array_equal(getattr(w, p + "t"), g[f"pt_{p}t"])
    w.nh_final()
    for p in "lp":
        for f in "vo":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"nhfinal_{p}{f}"], err_msg=p + f)
    w.nh_update()
    assert np.float32(w.zeta) == g["nhupd_zeta"] and np.float32(w.Q.value) == g["nhupd_Q"]
    assert int(g["hits_bounce_plain"]) > 20 and int(g["hits_bounce_fused"]) > 20
    w.box = (-float(g["bb_box"]), float(g["bb_box"]))
    w.bounce_back()
    for p in "lp":
        for f in "xv":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"bb_{p}{f}"], err_msg=p + f)
    w.box = (-float(g["nhi_box"]), float(g["nhi_box"]))
    w.nh_initial_fused()
    assert np.float32(w.zeta) == g["nhi_zeta"]
    for p in "lp":
        for f in "xvnoft":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"nhi_{p}{f}"], err_msg=p + f)
    # Langevin step with the noise the reference drew (stored in the fixture)
    w = port.World(sub(g, "ln_in_"), kBT=0.22)
    w.lf, w.lt, w.pf, w.pt = (g[k].copy() for k in ("ln_in_lf", "ln_in_lt", "ln_in_pf", "ln_in_pt"))
    w.verlet_langevin(g["ln_noise_l"], g["ln_noise_p"])
    for p in "lp":
        for f in "xvno":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"ln_{p}{f}"], err_msg=p + f)


def test_delete_lipid_with_survivors():
    g = load("branches_delete")
    w = port.World(sub(g, "del_in_"), kBT=0.0)
    n0 = len(w.lx)
    assert w.delete_lipid(2.5) == int(g["del_n"]) < n0
    for f in "xvno":
        np.testing.assert_array_equal(getattr(w, "l" + f), g[f"del_l{f}"], err_msg=f)
    np.testing.assert_array_equal(w.cs_l, g["del_cs_l"])
    assert w.delete_lipid(2.5) == int(g["del_n"])                  # nothing left to delete
    assert w.delete_lipid(1.2) == int(g["del2_n"]) < int(g["del_n"])
    for f in "xvno":
        np.testing.assert_array_equal(getattr(w, "l" + f), g[f"del2_l{f}"], err_msg=f)
    np.testing.assert_array_equal(w.cs_l, g["del2_cs_l"])
