"""The C restatement (oracle/orbc_oracle.c) against the committed golden vectors in tests/golden/*.npz, which were
produced by the reference itself (tests/golden/make_golden.py).  Runs anywhere: needs neither the reference tree nor a GPU."""
import os

import numpy as np
import pytest

from oracle import port
from tests.common import GOLDEN, rel_err

FILES = ["vesicle_ico0", "sphere_r12"]


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.mark.parametrize("name", FILES)
def test_chain(name):
    g = load(name)
    assert g["forcefield"].tobytes() == port.forcefield().as_array().tobytes()
    w = port.World({k[3:]: v for k, v in g.items() if k.startswith("in_")}, kBT=0.0)
    w.compute_pairwise_fused()
    for p in "lp":
        for f in "ft":
            assert rel_err(getattr(w, p + f), g[f"pair_{p}{f}"]) < 2e-5, (p, f)
    # continue from the reference's own pair forces so that later stages can be compared bit for bit
    w.lf, w.lt, w.pf, w.pt = (g[k].copy() for k in ("pair_lf", "pair_lt", "pair_pf", "pair_pt"))
    w.compute_bonded()
    np.testing.assert_array_equal(w.pf, g["bonded_pf"])
    assert w.compute_temperature() == float(g["temperature"])
    w.verlet_langevin()
    for p in "lp":
        for f in "xvno":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"lang_{p}{f}"], err_msg=p + f)
    # rebuild with Morton sort
    w.nstep = 24
    w.rebuild()
    np.testing.assert_array_equal(w.centroids, g["rb_centroids"])
    for p in "lp":
        aff, tie = getattr(w, "aff_" + p), getattr(w, "tie_" + p)
        bad = aff != g[f"rb_aff_{p}"]
        assert not (bad & ~tie).any()
        assert not bad.any(), "fixture chosen without ties"
        np.testing.assert_array_equal(getattr(w, "cs_" + p), g[f"rb_cs_{p}"])
        np.testing.assert_array_equal(getattr(w, "cells_" + p), g[f"rb_cells_{p}"])
        for f in "xvno":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"rb_{p}{f}"])
    np.testing.assert_array_equal(w.ptype, g["rb_ptype"]); np.testing.assert_array_equal(w.ptag, g["rb_ptag"])
    for k in (9, 8, 6):
        ptr, idx = g[f"st{k}_ptr"], g[f"st{k}_idx"]
        for c in range(w.n_cells):
            np.testing.assert_array_equal(port.stencil(w.centroids, c, float(k)), idx[ptr[c]:ptr[c + 1]])
    w.clear_force(); w.compute_pairwise_fused(); w.compute_bonded()
    for p in "lp":
        for f in "ft":
            assert rel_err(getattr(w, p + f), g[f"f2_{p}{f}"]) < 2e-5, (p, f)
    w.lf, w.lt, w.pf, w.pt = (g[k].copy() for k in ("f2_lf", "f2_lt", "f2_pf", "f2_pt"))
    w.kBT = 0.22; w.zeta = 0.03
    w.nh_final_fused()
    assert np.float32(w.zeta) == g["nhf_zeta"] and np.float32(w.Q.value) == g["nhf_Q"]
    for p in "lp":
        for f in "vot":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"nhf_{p}{f}"])
    w.nh_initial_fused()
    assert np.float32(w.zeta) == g["nhi_zeta"]
    for p in "lp":
        for f in "xvnoft":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"nhi_{p}{f}"])
    assert w.compute_temperature() == float(g["nhi_temperature"])


@pytest.mark.parametrize("name", FILES)
def test_keys_and_rng(name):
    g = load(name)
    keys = np.array([port.morton_encode(*p) for p in g["rb_centroids"]], np.uint32)
    # rb_centroids were permuted again by later calls? no: morton_keys were taken from the final centroids == rb_centroids
    np.testing.assert_array_equal(keys, g["morton_keys"])
    assert (np.diff(keys.astype(np.int64)) > 0).all()  # sorted, no duplicates
    out = np.array([port.lib().orc_uint2u11(int(u)) for u in g["u2u11_in"]], np.float32)
    np.testing.assert_array_equal(out, g["u2u11_out"])


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 distribution)."""
    assert port.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert port.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_noise_statistics():
    z = port.philox_noise(1234, 7, 0, 200000)
    assert z.min() >= -1.0 and z.max() <= 1.0
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0 / 3.0) < 5e-3
    z2 = port.philox_noise(1234, 8, 0, 200000)
    assert abs(np.corrcoef(z[:, 0], z2[:, 0])[0, 1]) < 1e-2


# ---- second set: frames, minimisation loop, volume constraint (tests/golden/make_golden_ext.py) ------------------------------------
def load_ext(name):
    return dict(np.load(os.path.join(GOLDEN, "ext_" + name + ".npz")))


def ext_state(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("in_")}


@pytest.mark.parametrize("name", FILES)
def test_frame_bytes(name):
    """save_frame's byte layout restated (oracle/port.py frame_bytes) == the reference's own file, byte for byte."""
    g = load_ext(name)
    st = ext_state(g)
    aff = [np.repeat(np.arange(len(cs) - 1, dtype=np.int32), np.diff(cs)) for cs in (st["cs_l"], st["cs_p"])]
    for df in (7, 31, 1):
        mine = port.frame_bytes(int(g["frame_nstep"]), df, int(g["tag_base"]), st["lx"], st["lv"], st["ln"], g["frc_lf"], aff[0],
                                st["px"], st["pv"], st["pn"], g["frc_pf"], aff[1], st["ptype"], st["ptag"])
        assert mine == g[f"frame_{df}"].tobytes(), df


@pytest.mark.parametrize("name", FILES)
def test_constrain_volume_golden(name):
    """Two consecutive calls (the scratch normals persist between them, constrain_volume.h:34,55) from a zeroed scratch."""
    g = load_ext(name)
    w = port.World(ext_state(g), kBT=0.0)
    for k in (1, 2):
        w.clear_force()
        vol = w.constrain_volume(3.15, 0.05)
        assert np.isfinite(vol)
        assert rel_err(w.lf, g[f"cv{k}_lf"]) < 1e-5, k
        assert rel_err(w.pf, g[f"cv{k}_pf"]) < 1e-5, k


@pytest.mark.parametrize("name", FILES)
def test_minimisation_loop_golden(name):
    """Two iterations of openrbc.cpp:88-133 (rebuild with Morton sort, clear, forces, post_torque, mover, bounce_back)."""
    g = load_ext(name)
    w = port.World(ext_state(g), kBT=0.0)
    for _ in range(2):
        w.nstep = 0
        w.rebuild()
        w.clear_force(); w.compute_pairwise_fused(); w.compute_bonded()
        w.post_torque(); w.opt_move(); w.bounce_back()
    np.testing.assert_array_equal(w.cs_l, g["opt_cs_l"])
    np.testing.assert_array_equal(w.cs_p, g["opt_cs_p"])
    # positions carry the summation-order ulps of the pair forces (stencil visiting order), so do their per-cell means
    assert rel_err(w.centroids, g["opt_centroids"]) < 1e-6
    for p in "lp":
        for f in "xn":
            assert rel_err(getattr(w, p + f), g[f"opt_{p}{f}"]) < 1e-6, (p, f)
        assert rel_err(getattr(w, p + "f"), g[f"opt_{p}f"]) < 2e-5


@pytest.mark.parametrize("name", ["sphere_r12", "sphere_r16"])
def test_voronoi_init_golden(name):
    """VoronoiDiagram::init (voronoi.h:54-75) restated == the reference's own result, bit for bit."""
    g = dict(np.load(os.path.join(GOLDEN, "init_" + name + ".npz")))
    c, cs, x, (n,), ties = port.voronoi_init(g["x0"], int(g["n_cells"]), int(g["n_iter"]), others=(g["n0"],))
    assert ties == 0
    np.testing.assert_array_equal(c, g["centroids"])
    np.testing.assert_array_equal(cs, g["cs_l"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(n, g["n"])
