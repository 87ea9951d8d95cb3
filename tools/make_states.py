"""Generates benchmark / large-test input states with the reference's own host-side initialisation
(init_rbc.h / init_random.h / VoronoiDiagram::init — code the north star leaves in place) through
oracle/_ref/libref_fast.so, followed by `--opt` steps of the reference's energy minimisation
(openrbc.cpp:88-146) so the membrane is relaxed like after `-E <n>`.

Run HERE (needs /root/reference for the mesh; the GPU box only sees the files this writes):
    python tools/make_states.py rbc --opt 100
    python tools/make_states.py sphere --radius 100
Outputs data/_gen/<name>.npz (git-ignored, travels with gpurun).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as refmod  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["rbc", "sphere", "ico"])
ap.add_argument("--opt", type=int, default=100)
ap.add_argument("--radius", type=float, default=100.0)
ap.add_argument("--subdiv", type=int, default=3)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--out", default=None)
a = ap.parse_args()

t0 = time.time()
if a.what == "rbc":
    mesh = os.path.join(ROOT, "oracle", "_ref", "example-large", "rbc")
    r = refmod.Ref("fast", threads=a.threads, args=["-i", "trimesh", "-m", mesh])
    r.init_trimesh()
    name = "rbc"
elif a.what == "ico":
    from tests.common import mesh_prefix
    r = refmod.Ref("fast", threads=a.threads, args=["-i", "trimesh", "-m", mesh_prefix(a.subdiv)])
    r.init_trimesh()
    name = f"ico{a.subdiv}"
else:
    r = refmod.Ref("fast", threads=a.threads, args=["-i", "lipid"])
    r.init_lipid_sphere(a.radius)
    name = f"sphere{int(a.radius)}"
print("init", r.size(0), r.size(1), f"{time.time() - t0:.1f}s", flush=True)
r.voronoi_init(64)
print("voronoi", r.n_cells, f"{time.time() - t0:.1f}s", flush=True)
if a.opt:
    r.set_param("stray_tolerance", 1e9)
    sec = r.run_opt(a.opt)
    print(f"opt {a.opt} steps {sec:.1f}s", flush=True)
    # leave the state partitioned consistently with its centroids (openrbc.cpp:155-157)
    r.voronoi_update(); r.cell_update(0); r.cell_update(1)
st = r.state()
st["lipid_tag_base"] = np.int32(r.lipid_tag_base())
st["opt_steps"] = np.int32(a.opt)
out = a.out or os.path.join(ROOT, "data", "_gen", name + ".npz")
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez(out, **st)
print("wrote", out, os.path.getsize(out) >> 20, "MiB", f"{time.time() - t0:.1f}s")
