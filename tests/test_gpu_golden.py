"""CUDA path (through the C ABI, openrbc_b200.Simulation) against the golden vectors produced by the reference itself
(tests/golden/*.npz, generator tests/golden/make_golden.py).  Teacher-forced: every stage starts from the reference's own
output of the previous stage, so integer structures can be required to be EQUAL and fp32 fields compared tightly.

Tolerances: forces / torques 1e-4 * (|f_ref| + f_rms) per particle; integrator outputs 1e-6
relative (1e-5 for the torque cross product n x t, which cancels); kinetic energy / temperature 1e-6 relative (fp64
accumulation on both sides)."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

FILES = ["vesicle_ico0", "sphere_r12"]
F_TOL, X_TOL, T_TOL = 1e-4, 1e-6, 1e-5


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def state_of(g, prefix):
    st = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    if prefix != "in":
        for p in "lp":
            for f in "xvno":
                st[p + f] = g[f"{prefix}_{p}{f}"]
    return st


@pytest.mark.parametrize("name", FILES)
def test_forces_on_input_state(name):
    from openrbc_b200 import Simulation
    g = load(name)
    sim = Simulation(state_of(g, "in"), kBT=0.0)
    sim.compute_pairwise_fused()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "ft")
        assert rel_err(d["f"], g[f"pair_{p}f"]) < F_TOL, (p, "f")
        assert rel_err(d["t"], g[f"pair_{p}t"]) < F_TOL, (p, "t")
    # teacher-force the reference's pair forces, then bonds
    for s, p in ((0, "l"), (1, "p")):
        sim.set_field(s, "f", g[f"pair_{p}f"]); sim.set_field(s, "t", g[f"pair_{p}t"])
    sim.compute_bonded()
    assert rel_err(sim.get(1, "f"), g["bonded_pf"]) < 2e-6
    assert abs(sim.compute_temperature() - float(g["temperature"])) <= 1e-6 * abs(float(g["temperature"])) + 1e-300
    # noise-free Langevin step from the reference's forces
    sim.set_field(1, "f", g["bonded_pf"])
    sim.verlet_langevin()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xvnoft")
        for f in "xvno":
            assert rel_err(d[f], g[f"lang_{p}{f}"]) < X_TOL, (p, f)
        assert not d["f"].any() and not d["t"].any()
    sim.close()


@pytest.mark.parametrize("name", FILES)
def test_rebuild_integer_exact(name):
    """voronoi.update + cell_lipid.update + cell_protein.update at nstep = 24 (Morton sort of the centroids included)."""
    from openrbc_b200 import Simulation
    g = load(name)
    sim = Simulation(state_of(g, "lang"), kBT=0.0)      # bit-identical input: the reference's post-Langevin state
    sim.nstep = 24
    sim.rebuild()
    np.testing.assert_array_equal(sim.dump("centroids"), g["rb_centroids"])
    np.testing.assert_array_equal(sim.dump("morton_keys"), g["morton_keys"])
    for s, p in ((0, "l"), (1, "p")):
        np.testing.assert_array_equal(sim.dump("aff_" + p), g[f"rb_aff_{p}"])
        np.testing.assert_array_equal(sim.dump("cell_start_" + p), g[f"rb_cs_{p}"])
        np.testing.assert_array_equal(sim.dump("cells_" + p), g[f"rb_cells_{p}"])
        d = sim.download(s, "xvno", ids=True, affiliation=True)
        for f in "xvno":
            np.testing.assert_array_equal(d[f], g[f"rb_{p}{f}"], err_msg=p + f)
        cs = g[f"rb_cs_{p}"]
        np.testing.assert_array_equal(d["affiliation"], np.repeat(np.arange(len(cs) - 1), np.diff(cs)))
        if s == 1:
            np.testing.assert_array_equal(d["type"], g["rb_ptype"]); np.testing.assert_array_equal(d["tag"], g["rb_ptag"])
    for k, got in zip((9, 8, 6), sim.stencils()):
        ptr, idx = g[f"st{k}_ptr"], g[f"st{k}_idx"]
        for c in range(sim.n_cells):
            np.testing.assert_array_equal(got[c], idx[ptr[c]:ptr[c + 1]], err_msg=f"stencil r<{k} cell {c}")
    # the stencil-guided search resolves (nearly) every particle; the exact grid search is only the fallback
    assert sim.dump("counters")[0] <= 0.05 * (sim.size(0) + sim.size(1))
    # forces on the rebuilt state
    sim.compute_pairwise_fused(); sim.compute_bonded()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "ft")
        assert rel_err(d["f"], g[f"f2_{p}f"]) < F_TOL and rel_err(d["t"], g[f"f2_{p}t"]) < F_TOL
    sim.close()


@pytest.mark.parametrize("name", FILES)
def test_nose_hoover_pair(name):
    from openrbc_b200 import Simulation
    g = load(name)
    st = state_of(g, "rb")
    st["ptype"], st["ptag"] = g["rb_ptype"], g["rb_ptag"]
    st["centroids"], st["cs_l"], st["cs_p"] = g["rb_centroids"], g["rb_cs_l"], g["rb_cs_p"]
    sim = Simulation(st, kBT=0.22)
    sim.zeta = 0.03
    for s, p in ((0, "l"), (1, "p")):
        sim.set_field(s, "f", g[f"f2_{p}f"]); sim.set_field(s, "t", g[f"f2_{p}t"])
    sim.nh_final_fused()
    assert abs(sim.zeta - float(g["nhf_zeta"])) <= 1e-6 * abs(float(g["nhf_zeta"]))
    assert sim.Q.value == float(g["nhf_Q"])
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "vot")
        for f in "vot":
            # t = n x t is a difference of products: the device contracts it into FMAs, the strict reference build does not,
            # so the cancellation leaves a few fp32 ulps of the operands rather than of the result
            assert rel_err(d[f], g[f"nhf_{p}{f}"]) < (T_TOL if f == "t" else X_TOL), (p, f)
    # second kernel from the reference's own intermediate state
    for s, p in ((0, "l"), (1, "p")):
        for f in "vot":
            sim.set_field(s, f, g[f"nhf_{p}{f}"])
    sim.zeta = float(g["nhf_zeta"])
    sim.nh_initial_fused()
    assert abs(sim.zeta - float(g["nhi_zeta"])) <= 1e-6 * abs(float(g["nhi_zeta"]))
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "xvnoft")
        for f in "xvno":
            assert rel_err(d[f], g[f"nhi_{p}{f}"]) < X_TOL, (p, f)
        assert not d["f"].any() and not d["t"].any()
    t_ref = float(g["nhi_temperature"])
    assert abs(sim.compute_temperature() - t_ref) <= 1e-6 * t_ref
    sim.close()


def test_counter_based_rng_matches_port_and_law():
    """Philox4x32-10 on the device == the port's restatement (itself checked against Random123 known answers), and the
    noise has rng.h's law: uniform in [-1, 1) with variance 1/3 (the sqrt(3) in sigma, integrate_langevin.h:113)."""
    from openrbc_b200 import Simulation
    from oracle import port
    sim = Simulation(None, seed=0x1234ABCD5678)
    z = sim.noise(7, 1, 100000)
    np.testing.assert_array_equal(z, port.philox_noise(0x1234ABCD5678, 7, 1, 100000))
    assert z.min() >= -1.0 and z.max() <= 1.0
    assert abs(z.mean()) < 1e-2 and abs(z.var() - 1 / 3) < 1e-2
    sim.close()


@pytest.mark.parametrize("name", FILES)
def test_tile_kernel_against_the_run_list_kernel(name):
    """k_pair_ll_t (warp per cell, shared-memory tile, prefilter + exact re-test) finds exactly the hits of k_pair_ll_r (thread per
    lipid): the forces differ only by the order of the per-lipid sums.  A cell whose candidates do not fit the tile sends the step
    to k_pair_ll_r (device flag): with the capacity shrunk by the test option the results are then bit-identical to that kernel's."""
    from openrbc_b200 import Simulation
    g = load(name)
    st = state_of(g, "in")
    out = {}
    for key, opts in (("tile", {"ll_variant": 0}), ("runs", {"ll_variant": 1}), ("overflow", {"ll_variant": 0, "debug_tile_cap": 48})):
        sim = Simulation(st, kBT=0.0)
        for k, v in opts.items():
            sim.set_option(k, v)
        sim.compute_pairwise_fused()
        out[key] = sim.download(0, "ft")
        # a second evaluation accumulates (reference semantics: f, t +=): exactly twice the first for a deterministic kernel
        if len(st["px"]) == 0:
            sim.compute_pairwise_fused()
            again = sim.download(0, "ft")
            np.testing.assert_array_equal(again["f"], 2 * out[key]["f"])
        sim.close()
    exact = len(st["px"]) == 0            # protein -> lipid reactions arrive by atomics: their order is not fixed
    for k in "ft":
        assert rel_err(out["tile"][k], out["runs"][k]) < 2e-6
        for other in ("overflow",):                          # same kernel, same order of the sums
            if exact:
                np.testing.assert_array_equal(out[other][k], out["runs"][k])
            else:
                assert rel_err(out[other][k], out["runs"][k]) < 1e-6
    # the tile kernel is deterministic: two fresh contexts give the same bits
    if exact:
        sim = Simulation(st, kBT=0.0); sim.set_option("ll_variant", 0); sim.compute_pairwise_fused()
        np.testing.assert_array_equal(sim.get(0, "f"), out["tile"]["f"])
        sim.close()
