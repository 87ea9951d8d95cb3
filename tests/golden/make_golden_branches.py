"""Third set of golden vectors, produced by the REFERENCE ITSELF (oracle/_ref/libref_strict.so, one thread): states built so
that every branch of the pair driver, the wall reflection and delete_lipid actually FIRE.   python tests/golden/make_golden_branches.py

The plain vesicle fixtures never reach the Lennard-Jones `else` branches (pairwise_kernel_fused.h:63-77), the protein-protein
repulsion (:79-97), the reflecting wall (integrate_nh.h:124-144,200-209) or a deletion with survivors < N (cleanup.h:29-91).
Here proteins of the vesicle are moved by hand next to lipids / next to each other, a few particles are put at the wall with
an outward velocity, and a few lipids are lifted off the membrane.  tests/common.py:branch_hits() counts, per branch, the pairs
the reference's driver evaluates (stencil-aware brute force); the generator and the tests assert that every count is > 0.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as refmod  # noqa: E402
from tests.common import branch_hits, ref_vesicle  # noqa: E402
from tests.golden.make_golden import snap, stencils_csr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def unit(v):
    return v / np.linalg.norm(v)


def tangent(n):
    a = np.array([1.0, 0.0, 0.0]) if abs(n[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    return unit(np.cross(n, a))


def perturb(st, rng):
    """Move proteins so that every pair branch fires.  Returns the new state (not yet partitioned consistently)."""
    st = {k: v.copy() for k, v in st.items()}
    lx, ln, px, ty = st["lx"], st["ln"], st["px"], st["ptype"]
    by_type = {t: list(np.flatnonzero(ty == t)) for t in range(6)}
    # a third of the band-3 become the mobile kind (type 1), so that type12 = type1 + 6 type2 is exercised on {1,2,3}^2
    for i in by_type[2][::3]:
        ty[i] = 1
    by_type = {t: list(np.flatnonzero(ty == t)) for t in range(6)}
    # sites: lipids at least 6 apart from each other, so that the hand-placed groups never see one another (the longest range
    # involved is the sigma = 3.4 LJ of band-3 against actin / spectrin, cut 3.8165).  The directors of this vesicle point INWARD
    # and its cytoskeleton lies outside the bilayer: everything is placed on the inner side, +n, where no other protein is.
    lip = []
    for j in rng.permutation(len(lx)):
        if all(np.linalg.norm(lx[j] - lx[k]) > 6.0 for k in lip):
            lip.append(j)
    assert np.dot(lx[lip[0]], ln[lip[0]]) < 0

    def lipid_site():
        j = lip.pop()
        return lx[j].astype(np.float64), unit(ln[j].astype(np.float64))

    # (a) protein-lipid LJ: spectrin (5) and actin (4) 0.97 .. 1.10 above a lipid (lj cut 1.1225, forcefield_canonical.h:92-99)
    for t, cnt in ((5, 20), (4, 4)):
        for _ in range(cnt):
            i = by_type[t].pop()
            x, n = lipid_site()
            px[i] = x + n * rng.uniform(0.97, 1.10) + tangent(n) * rng.uniform(-0.1, 0.1)
    # (b) protein-protein repulsion on {1,2,3}^2 (cut 2.6): pairs 1.9 .. 2.4 apart, 1.5 off the membrane
    pairs = [(1, 1), (1, 2), (2, 2), (1, 3), (2, 3), (3, 3)]
    for ta, tb in pairs * 2:
        i, j = by_type[ta].pop(), by_type[tb].pop()
        x, n = lipid_site()
        px[i] = x + n * 1.5
        px[j] = px[i] + tangent(n) * rng.uniform(1.9, 2.4)
    # (c) protein-protein LJ: {band-3 (1, 2)} x {actin (4), spectrin (5)}; sigma 3.4 (cut 3.8165) except (2, 5): sigma 1 (cut 1.1225)
    for ta, tb, lo, hi in ((1, 4, 3.0, 3.6), (1, 5, 3.0, 3.6), (2, 4, 3.0, 3.6), (2, 5, 0.98, 1.10), (2, 4, 3.2, 3.7), (2, 5, 1.0, 1.08)):
        i, j = by_type[ta].pop(), by_type[tb].pop()
        x, n = lipid_site()
        px[i] = x + n * 1.6
        px[j] = px[i] + unit(n + 0.3 * tangent(n)) * rng.uniform(lo, hi)
    return st


def forces_chain():
    r = ref_vesicle(0)
    r.set_param("kBT", 0.0)
    st = perturb(r.state(), np.random.default_rng(20261017))
    r.load_state(st)
    # a consistent partition of the perturbed state: the driver's own rebuild (Morton step included)
    r.set_param("nstep", 24)
    r.voronoi_update(); r.cell_update(0); r.cell_update(1)
    g = {}
    for k, v in r.state().items():
        g["in_" + k] = v
    for k, (ptr, idx) in zip((9, 8, 6), stencils_csr(r)):
        g[f"st{k}_ptr"], g[f"st{k}_idx"] = ptr, idx
    hits = branch_hits({k[3:]: v for k, v in g.items() if k.startswith("in_")}, g, r.forcefield())
    print("branch hits:", hits)
    assert all(v > 0 for v in hits.values()), hits
    for k, v in hits.items():
        g["hits_" + k] = np.int64(v)
    g["forcefield"] = r.forcefield()
    r.integrate(refmod.CLEAR_FORCE)
    r.compute_pairwise_fused(); snap(r, g, "pair", "ft")
    r.compute_bonded(); g["bonded_pf"] = r.get(1, "f")
    # one noise-free Langevin step and one rebuild from there (the perturbed proteins sit in unusual places)
    r.integrate(refmod.VERLET_LANGEVIN); snap(r, g, "lang", "xvno")
    r.set_param("nstep", 2)
    r.voronoi_update(); g["rb_centroids"] = r.centroids()
    for s, p in ((0, "l"), (1, "p")):
        r.cell_update(s)
        g[f"rb_aff_{p}"] = r.cell_array(s, "affiliation"); g[f"rb_cs_{p}"] = r.cell_array(s, "cell_start"); g[f"rb_cells_{p}"] = r.cell_array(s, "cells")
    g["rb_ptype"], g["rb_ptag"] = r.protein_ids()
    return r, g


def integrator_chain(r, g):
    """Unfused NH kernels, the wall, injected-noise Langevin.  Continues from the rebuilt state of forces_chain."""
    rng = np.random.default_rng(7)
    for s in (0, 1):
        r.set(s, "v", rng.normal(0, 0.4, (r.size(s), 3)))
        r.set(s, "o", rng.normal(0, 0.2, (r.size(s), 3)))
    r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded()
    r.set_param("kBT", 0.22); r.set_param("zeta", 0.04); r.set_param("Q", 0.0)
    st = r.state()
    for k, v in st.items():
        g["nh_in_" + k] = v
    snap(r, g, "nh_in", "ft")
    # --- unfused: post_torque, verlet_nh_final, verlet_nh_update (integrate_nh.h:146-176, 66-94)
    r.integrate(refmod.POST_TORQUE); snap(r, g, "pt", "t")
    r.integrate(refmod.NH_FINAL); snap(r, g, "nhfinal", "vo")
    r.integrate(refmod.NH_UPDATE); g["nhupd_zeta"] = np.float32(r.get_param("zeta")); g["nhupd_Q"] = np.float32(r.get_param("Q"))
    # --- the reflecting walls (integrate_nh.h:124-144): the box is drawn THROUGH the vesicle (radius 14.5), so that the caps beyond
    #     +-14.3 are folded back on all six faces; the reference's box is +-1000 and nothing ever reaches it in the other fixtures
    g["bb_box"] = np.float64(14.3)
    r.set_param("box_lo", -14.3); r.set_param("box_hi", 14.3)
    r.integrate(refmod.BOUNCE_BACK); snap(r, g, "bb", "xv")
    g["hits_bounce_plain"] = np.int64(sum(int((np.abs(g[f"nh_in_{p}x"]) > 14.3).sum()) for p in "lp"))
    for p in "lp":
        moved = (g[f"bb_{p}x"] != g[f"nh_in_{p}x"])
        assert (moved == (np.abs(g[f"nh_in_{p}x"]) > 14.3)).all()
    # --- fused initial kernel with a tighter box: drift, reflection (integrate_nh.h:200-209), kinetic energy, zeta
    g["nhi_box"] = np.float64(14.2)
    r.set_param("box_lo", -14.2); r.set_param("box_hi", 14.2)
    r.integrate(refmod.NH_INITIAL_FUSED); snap(r, g, "nhi", "xvnoft"); g["nhi_zeta"] = np.float32(r.get_param("zeta"))
    g["hits_bounce_fused"] = np.int64(sum(int((np.abs(g[f"bb_{p}x"]) > 14.21).sum()) for p in "lp"))
    faces = sum(int(((g[f"bb_{p}x"] > 14.21).any(0)).sum() + ((g[f"bb_{p}x"] < -14.21).any(0)).sum()) for p in "l")
    print("bounce hits: plain", int(g["hits_bounce_plain"]), "fused", int(g["hits_bounce_fused"]), "faces hit by lipids", faces)
    assert g["hits_bounce_plain"] > 20 and g["hits_bounce_fused"] > 20 and faces == 6
    r.set_param("box_lo", -1000.0); r.set_param("box_hi", 1000.0)
    # --- one Langevin step with the reference's own noise (1 thread: MT19937 -> xorshift128 stream, integrate_langevin.h:116-137);
    #     the noise itself is recovered by the port from the seed (tests/test_oracle_vs_ref.py pins that restatement), so the
    #     fixture stores it for the device's injection hook
    from oracle import port
    seed = 0xBAD5EED
    mt0 = port.mt_init(seed)
    prng0 = port.mt_init(port.lib().orc_mt_uint(port.C.byref(mt0)))
    for s in (0, 1):                                             # the noise-free Langevin step of forces_chain drew from the same stream
        port.langevin_noise(prng0, r.size(s))
    r.load_state(st)                                             # back to the state before the walls folded particles onto each other
    r.set_param("box_lo", -1000.0); r.set_param("box_hi", 1000.0)
    r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded()
    for k, v in r.state().items():
        g["ln_in_" + k] = v
    snap(r, g, "ln_in", "ft")
    g["ln_noise_l"] = port.langevin_noise(prng0, r.size(0)); g["ln_noise_p"] = port.langevin_noise(prng0, r.size(1))
    r.integrate(refmod.VERLET_LANGEVIN); snap(r, g, "ln", "xvno")
    # the stored noise really is what the reference drew: the port's step from it reproduces the reference bit for bit
    w = port.World({k[6:]: v for k, v in g.items() if k.startswith("ln_in_") and k[6:] not in ("lf", "lt", "pf", "pt")}, kBT=0.22)
    w.lf, w.lt, w.pf, w.pt = g["ln_in_lf"].copy(), g["ln_in_lt"].copy(), g["ln_in_pf"].copy(), g["ln_in_pt"].copy()
    w.verlet_langevin(g["ln_noise_l"], g["ln_noise_p"])
    for p in "lp":
        for f in "xvno":
            np.testing.assert_array_equal(getattr(w, p + f), g[f"ln_{p}{f}"], err_msg=p + f)


def delete_chain():
    """delete_lipid with survivors < N (cleanup.h:29-91): lipids lifted off the membrane by more than the stray tolerance."""
    r = ref_vesicle(0)
    r.set_param("kBT", 0.0)
    rng = np.random.default_rng(99)
    x, n = r.get(0, "x"), r.get(0, "n")
    cs = r.cell_array(0, "cell_start")
    lifted = []
    # two strays in one cell, one in others, one in a cell's FIRST and one in a cell's LAST slot; and a near miss that must survive
    cells = rng.permutation(r.n_cells)[:9]
    for k, c in enumerate(cells):
        b, e = cs[c], cs[c + 1]
        picks = [b] if k == 0 else [e - 1] if k == 1 else [b + 1, b + 2] if k == 2 else [b + (e - b) // 2]
        for j in picks:
            x[j] += unit(n[j]) * np.float32(7.0 + k)
            lifted.append(j)
    r.set(0, "x", x)
    r.set_param("stray_tolerance", 2.5)
    g = {}
    for k, v in r.state().items():
        g["del_in_" + k] = v
    g["del_lifted"] = np.array(sorted(lifted), np.int32)
    n0 = r.size(0)
    n1 = r.delete_lipid()
    g["del_n"] = np.int64(n1)
    assert 0 < n0 - n1 <= len(lifted), (n0, n1)
    print("delete_lipid:", n0, "->", n1, "lifted", len(lifted))
    snap(r, g, "del", "xvno")
    g["del_cs_l"] = r.cell_array(0, "cell_start"); g["del_aff_l"] = r.cell_array(0, "affiliation")
    # second call: nothing left to delete at this tolerance -> size unchanged, arrays untouched
    assert r.delete_lipid() == n1
    # and a tight tolerance that removes a good part of every cell
    r.set_param("stray_tolerance", 1.2)
    n2 = r.delete_lipid()
    g["del2_n"] = np.int64(n2); snap(r, g, "del2", "xvno"); g["del2_cs_l"] = r.cell_array(0, "cell_start")
    print("delete_lipid (tolerance 1.2):", n1, "->", n2)
    assert n2 < n1
    return g


if __name__ == "__main__":
    import subprocess
    if len(sys.argv) > 1 and sys.argv[1] == "delete":
        g = delete_chain()
        np.savez_compressed(os.path.join(OUT, "branches_delete.npz"), **g)
    elif len(sys.argv) > 1:
        r, g = forces_chain()
        integrator_chain(r, g)
        path = os.path.join(OUT, "branches_vesicle_ico0.npz")
        np.savez_compressed(path, **g)
        print(os.path.getsize(path) // 1024, "KiB")
    else:
        for what in ("forces", "delete"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), what])
