"""Parity at the BENCHMARKED configuration: the full red blood cell of BASELINE.json configs[1] (3 205 506 particles, 188 549 Voronoi
cells, extent +-539 x +-539 x +-178, the state bench.py times), teacher-forced against the unmodified reference
(oracle/_ref/libref_strict.so, one thread, travels to the GPU box prebuilt).

One rebuild at nstep = 24 (centroid update, Morton sort of the centroids — the 2x + 2000 >= 2048 key seam at |x| ~ 24 lies well
inside this system, reorder_morton.h:37-42 —, nearest-centroid partition of both containers, reorder) and one force evaluation:
  * centroid bits, Morton keys, affiliation, cell_start, cells, the reordered x v n o, protein ids: EQUAL
  * r<9 / r<8 / r<6 centroid stencil sets of every cell: EQUAL
  * pair + bonded forces and torques: |d| <= 1e-4 (|f_ref| + f_rms) per particle
The run writes a summary to gpurun_out/rbc_scale_parity.json (copied to profiles/ by hand after a GPU session).
Takes about two minutes (the state is generated once by the reference's own initialisation, shared with bench.py)."""
import ctypes as C
import json
import os
import time

import numpy as np
import pytest

from tests.common import HAVE_REF, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_stencil_table(r, nc, stride=64):
    """(nc, 3, stride) sorted stencil ids padded with INT_MAX, and (nc, 3) counts, from the reference's tree search."""
    big = np.iinfo(np.int32).max
    tab = np.full((nc, 3, stride), big, np.int32)
    cnt = np.zeros((nc, 3), np.int32)
    o = [np.empty(256, np.int32) for _ in range(3)]
    n = np.zeros(3, np.int32)
    ptr = [a.ctypes.data_as(C.c_void_p) for a in o]
    nptr = n.ctypes.data_as(C.c_void_p)
    fn = r.lib.ref_get_stencil_refined
    for c in range(nc):
        fn(c, ptr[0], ptr[1], ptr[2], 256, nptr)
        cnt[c] = n
        for k in range(3):
            tab[c, k, :n[k]] = o[k][:n[k]]
    tab.sort(axis=2)
    return tab, cnt


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_strict.so not built")
def test_full_rbc_rebuild_and_forces_match_the_reference():
    import bench
    from openrbc_b200 import Simulation
    from oracle import ref as refmod
    t0 = time.time()
    st = bench.load_state("rbc")
    st = {k: st[k] for k in ("lx", "lv", "ln", "lo", "px", "pv", "pn", "po", "ptype", "ptag", "bonds", "centroids", "cs_l", "cs_p")}
    n_l, n_p, nc = len(st["lx"]), len(st["px"]), len(st["centroids"])
    assert n_l + n_p > 3_000_000 and nc > 150_000
    assert np.abs(st["lx"]).max() > 500.0                          # spans the Morton seam and then some
    log = {"particles": n_l + n_p, "lipids": n_l, "proteins": n_p, "cells": nc, "bonds": len(st["bonds"]), "extent": np.abs(st["lx"]).max(0).round(1).tolist()}

    r = refmod.Ref("strict", threads=1, args=["-i", "lipid"])
    r.load_state(st)
    r.set_param("kBT", 0.0); r.set_param("nstep", 24)
    r.voronoi_update(); r.cell_update(0); r.cell_update(1)
    sim = Simulation(st, kBT=0.0)
    sim.nstep = 24
    sim.rebuild()
    sim.synchronize()                                              # raises if a stencil overflowed (n6 > 32 or more than 64 cells)
    log["t_setup_s"] = round(time.time() - t0, 1)

    cen = r.centroids()
    np.testing.assert_array_equal(sim.dump("centroids"), cen)
    keys = sim.dump("morton_keys")
    enc = r.lib.ref_morton_encode
    ref_keys = np.fromiter((enc(float(p[0]), float(p[1]), float(p[2])) for p in cen), np.uint32, nc)
    np.testing.assert_array_equal(keys, ref_keys)
    assert (np.diff(keys.astype(np.int64)) > 0).all()              # sorted and free of duplicates: the stable sort is the reference's order
    seam = int(((2.0 * cen[:, 0] + 2000.0 >= 2048.0) & (2.0 * cen[:, 0] + 2000.0 < 2049.0)).sum())
    log["centroids_on_the_x_seam"] = seam
    assert seam > 0
    for s, p in ((0, "l"), (1, "p")):
        np.testing.assert_array_equal(sim.dump("aff_" + p), r.cell_array(s, "affiliation"), err_msg="affiliation " + p)
        cs = r.cell_array(s, "cell_start")
        np.testing.assert_array_equal(sim.dump("cell_start_" + p), cs, err_msg="cell_start " + p)
        np.testing.assert_array_equal(sim.dump("cells_" + p), r.cell_array(s, "cells"), err_msg="cells " + p)
        d = sim.download(s, "xvno", ids=(s == 1))
        for f in "xvno":
            np.testing.assert_array_equal(d[f], r.get(s, f), err_msg=p + f)
        if s == 1:
            ty, tg = r.protein_ids()
            np.testing.assert_array_equal(d["type"], ty); np.testing.assert_array_equal(d["tag"], tg)
        log[f"cell_size_{p}"] = [int(np.diff(cs).min()), round(float(np.diff(cs).mean()), 2), int(np.diff(cs).max())]
    # stencil sets of every cell
    tab, cnt = ref_stencil_table(r, nc)
    dcnt = sim.dump("stencil_counts")                             # (nc, 3): n6, n8, n9
    np.testing.assert_array_equal(dcnt[:, ::-1], cnt)             # reference order here: 9, 8, 6
    dst = sim.dump("stencil")
    big = np.iinfo(np.int32).max
    col = np.arange(dst.shape[1])[None, :]
    for k, name in ((0, 9), (1, 8), (2, 6)):
        mine = np.where(col < cnt[:, k][:, None], dst, big)
        mine.sort(axis=1)
        np.testing.assert_array_equal(mine, tab[:, k, :], err_msg=f"stencil r<{name}")
        log[f"stencil{name}"] = [int(cnt[:, k].min()), round(float(cnt[:, k].mean()), 2), int(cnt[:, k].max())]
    # forces on the rebuilt state
    r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded()
    sim.compute_pairwise_fused(); sim.compute_bonded()
    for s, p in ((0, "l"), (1, "p")):
        d = sim.download(s, "ft")
        for f in "ft":
            e = rel_err(d[f], r.get(s, f))
            log[f"rel_err_{p}{f}"] = e
            assert e < 1e-4, (p, f, e)
    # ---- a second rebuild two steps later, without renumbering: the device re-classifies the recorded wide stencils (k_stencil_refresh)
    #      instead of searching the grid; sets and partition must still be the reference's
    for s, p in ((0, "l"), (1, "p")):                              # the same (reference) forces on both sides, then one noise-free step
        sim.set_field(s, "f", r.get(s, "f")); sim.set_field(s, "t", r.get(s, "t"))
    r.integrate(refmod.VERLET_LANGEVIN); sim.verlet_langevin()
    for s, p in ((0, "l"), (1, "p")):                              # teacher-force the reference's positions (the integrators agree to 1e-6 only)
        for f in "xvno":
            sim.set_field(s, f, r.get(s, f))
    r.set_param("nstep", 26); sim.nstep = 26
    r.voronoi_update(); r.cell_update(0); r.cell_update(1)
    sim.rebuild(); sim.synchronize()
    np.testing.assert_array_equal(sim.dump("centroids"), r.centroids())
    for s, p in ((0, "l"), (1, "p")):
        np.testing.assert_array_equal(sim.dump("cell_start_" + p), r.cell_array(s, "cell_start"), err_msg="second rebuild cell_start " + p)
        np.testing.assert_array_equal(sim.dump("aff_" + p), r.cell_array(s, "affiliation"), err_msg="second rebuild affiliation " + p)
    tab, cnt = ref_stencil_table(r, nc)
    np.testing.assert_array_equal(sim.dump("stencil_counts")[:, ::-1], cnt)
    dst = sim.dump("stencil")
    changed = 0
    for k, name in ((0, 9), (1, 8), (2, 6)):
        mine = np.where(col < cnt[:, k][:, None], dst, big)
        mine.sort(axis=1)
        np.testing.assert_array_equal(mine, tab[:, k, :], err_msg=f"second rebuild stencil r<{name}")
    log["second_rebuild"] = "refresh path (k_stencil_refresh): stencil sets, affiliation and cell_start equal to the reference"
    log["t_total_s"] = round(time.time() - t0, 1)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "rbc_scale_parity.json"), "w") as fh:
        json.dump(log, fh, indent=1)
    print(json.dumps(log))
    sim.close()
