"""Second set of golden vectors, again produced by the REFERENCE ITSELF (oracle/_ref/libref_strict.so, one thread):
the trajectory frame bytes (trajectory.h:61-105), the energy-minimisation loop (openrbc.cpp:88-133) and
constrain_volume (constrain_volume.h:26-83).   python tests/golden/make_golden_ext.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as refmod  # noqa: E402
from tests.common import ref_sphere, ref_vesicle  # noqa: E402
from tests.golden.make_golden import snap  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def chain(r, name):
    g = {}
    r.set_param("kBT", 0.0)
    for _ in range(3):
        r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded(); r.integrate(refmod.VERLET_LANGEVIN)
    r.set_param("nstep", 24)
    r.voronoi_update(); r.cell_update(0); r.cell_update(1)
    for k, v in r.state().items():
        g["in_" + k] = v
    g["tag_base"] = np.int32(r.lipid_tag_base())
    # --- frames: forces present so that the FORCE section is not all zeros
    r.integrate(refmod.CLEAR_FORCE); r.compute_pairwise_fused(); r.compute_bonded()
    snap(r, g, "frc", "ft")
    g["frame_nstep"] = np.int32(24)
    for df in (7, 31, 1):
        g[f"frame_{df}"] = np.frombuffer(r.save_frame(df), np.uint8)
    # --- constrain_volume twice on cleared forces; the reference's scratch is uninitialised malloc memory, M_PERTURB makes the
    #     first call see zeros (what the device starts from)
    libc = ctypes.CDLL(None)
    for k in (1, 2):
        r.integrate(refmod.CLEAR_FORCE)
        libc.mallopt(-6, 255)
        try:
            r.constrain_volume(3.15, 0.05)
        finally:
            libc.mallopt(-6, 0)
        snap(r, g, f"cv{k}", "f")
        assert np.isfinite(g[f"cv{k}_lf"]).all()
    # --- two iterations of the minimisation loop from the input state (param.nstep = 0 there: Morton sort every time)
    r.set_param("nstep", 0)
    r.run_opt(2)
    snap(r, g, "opt", "xvnoft")
    g["opt_centroids"] = r.centroids()
    g["opt_cs_l"] = r.cell_array(0, "cell_start"); g["opt_cs_p"] = r.cell_array(1, "cell_start")
    g["opt_ptype"], g["opt_ptag"] = r.protein_ids()
    path = os.path.join(OUT, "ext_" + name + ".npz")
    np.savez_compressed(path, **g)
    print(name, os.path.getsize(path) // 1024, "KiB")


def init_chain(name, radius, n_iter):
    """VoronoiDiagram::init (voronoi.h:54-75) on a random lipid sphere: the container before, the Voronoi state and the container after."""
    r = refmod.Ref("strict", threads=1, args=["-i", "lipid"])
    r.init_lipid_sphere(radius)
    g = {"x0": r.get(0, "x"), "n0": r.get(0, "n"), "n_iter": np.int32(n_iter)}
    g["n_cells"] = np.int32(r.voronoi_init(n_iter))
    g["centroids"] = r.centroids(); g["cs_l"] = r.cell_array(0, "cell_start"); g["x"] = r.get(0, "x"); g["n"] = r.get(0, "n")
    path = os.path.join(OUT, "init_" + name + ".npz")
    np.savez_compressed(path, **g)
    print(name, len(g["x0"]), "lipids", int(g["n_cells"]), "cells", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    # one process per chain: constrain_volume's scratch is a function-static that would carry over from one world to the next
    import subprocess
    if len(sys.argv) > 1 and sys.argv[1].startswith("init_"):
        init_chain(sys.argv[1][5:], {"sphere_r12": 12.0, "sphere_r16": 16.0}[sys.argv[1][5:]], {"sphere_r12": 64, "sphere_r16": 11}[sys.argv[1][5:]])
    elif len(sys.argv) > 1:
        chain({"vesicle_ico0": lambda: ref_vesicle(0), "sphere_r12": lambda: ref_sphere(12.0)}[sys.argv[1]](), sys.argv[1])
    else:
        for name in ("vesicle_ico0", "sphere_r12", "init_sphere_r12", "init_sphere_r16"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
