"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_strict.so = unmodified
/root/reference/src headers, strict IEEE flags, one thread).  Run here (the reference tree is not on the
GPU box):   python tests/golden/make_golden.py
Each file holds one teacher-forced chain: input state -> outputs of every hot-path call.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as refmod  # noqa: E402
from tests.common import ref_sphere, ref_vesicle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def snap(r, g, prefix, fields="xvnoft"):
    for s, p in ((0, "l"), (1, "p")):
        for f in fields:
            g[f"{prefix}_{p}{f}"] = r.get(s, f)


def stencils_csr(r):
    ptr = [[0], [0], [0]]
    idx = [[], [], []]
    for c in range(r.n_cells):
        for k, s in enumerate(r.stencil_refined(c)):
            s = np.sort(s)
            idx[k].extend(s.tolist())
            ptr[k].append(len(idx[k]))
    return [(np.array(p, np.int32), np.array(i, np.int32)) for p, i in zip(ptr, idx)]


def chain(r, name):
    g = {}
    g["forcefield"] = r.forcefield()
    # relax a little so velocities / angular velocities are non-zero and the partition is stale
    r.set_param("kBT", 0.0)
    for _ in range(4):
        r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded(); r.integrate(refmod.VERLET_LANGEVIN)
    st = r.state()
    for k, v in st.items():
        g["in_" + k] = v
    # --- forces on the input state
    r.integrate(refmod.CLEAR_FORCE)
    r.compute_pairwise_fused(); snap(r, g, "pair", "ft")
    r.compute_bonded(); g["bonded_pf"] = r.get(1, "f")
    g["temperature"] = np.float64(r.compute_temperature())
    # --- Langevin step without noise
    r.integrate(refmod.VERLET_LANGEVIN); snap(r, g, "lang", "xvno")
    # --- rebuild with Morton sort (nstep % 24 == 0)
    r.set_param("nstep", 24)
    r.voronoi_update(); g["rb_centroids"] = r.centroids()
    for s, p in ((0, "l"), (1, "p")):
        r.cell_update(s)
        g[f"rb_aff_{p}"] = r.cell_array(s, "affiliation")
        g[f"rb_cs_{p}"] = r.cell_array(s, "cell_start")
        g[f"rb_cells_{p}"] = r.cell_array(s, "cells")
    snap(r, g, "rb", "xvno")
    g["rb_ptype"], g["rb_ptag"] = r.protein_ids()
    for k, (ptr, idx) in zip((9, 8, 6), stencils_csr(r)):
        g[f"st{k}_ptr"], g[f"st{k}_idx"] = ptr, idx
    # --- forces on the rebuilt state, then the Nose-Hoover pair
    r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded(); snap(r, g, "f2", "ft")
    r.set_param("kBT", 0.22); r.set_param("zeta", 0.03); r.set_param("Q", 0.0)
    r.integrate(refmod.NH_FINAL_FUSED); snap(r, g, "nhf", "vot"); g["nhf_zeta"] = np.float32(r.get_param("zeta")); g["nhf_Q"] = np.float32(r.get_param("Q"))
    r.integrate(refmod.NH_INITIAL_FUSED); snap(r, g, "nhi", "xvnoft"); g["nhi_zeta"] = np.float32(r.get_param("zeta"))
    g["nhi_temperature"] = np.float64(r.compute_temperature())
    # --- Morton keys of the centroids and a few RNG values
    c = r.centroids()
    g["morton_keys"] = np.array([r.morton_encode(*p) for p in c], np.uint32)
    u = np.array([0, 1, 2, 12345, 2**31 - 1, 2**31, 2**31 + 1, 2**32 - 2, 2**32 - 1], np.uint64)
    g["u2u11_in"] = u.astype(np.uint32); g["u2u11_out"] = np.array([r.uint2u11(int(x)) for x in u], np.float32)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **g)
    print(name, {k: v.shape for k, v in g.items() if hasattr(v, "shape") and k.startswith("in_")}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    chain(ref_vesicle(0), "vesicle_ico0")
    chain(ref_sphere(12.0), "sphere_r12")
