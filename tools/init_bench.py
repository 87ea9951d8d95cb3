"""VoronoiDiagram::init on the device (orbc_voronoi_init, 64 Lloyd rounds) on a workload's lipids: wall time and cell statistics.
    python tools/init_bench.py [workload]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

st = dict(bench.load_state(sys.argv[1] if len(sys.argv) > 1 else "rbc"))
nc = len(st.pop("centroids")); st.pop("cs_l"); st.pop("cs_p")
rng = np.random.default_rng(1)
perm = rng.permutation(len(st["lx"]))            # forget the order the reference's own init left behind
for k in ("lx", "lv", "ln", "lo"):
    st[k] = np.ascontiguousarray(st[k][perm])
sim = orbc.Simulation(st, kBT=0.22)
for rep in range(2):
    sim.upload(st)
    sim.synchronize()
    t0 = time.perf_counter()
    sim.voronoi_init(nc, 64)
    t1 = time.perf_counter()
    sim.cell_update(1)
    sim.synchronize()
    cs = sim.dump("cell_start_l")
    cnt = np.diff(cs)
    print(f"voronoi_init({nc} cells, 64 rounds) over {len(st['lx'])} lipids: {(t1 - t0) * 1e3:.1f} ms; lipids per cell min/mean/max {cnt.min()}/{cnt.mean():.1f}/{cnt.max()}; "
          f"grid fallback searches {int(sim.dump('counters')[0])}", flush=True)
sim.run_langevin(20); sim.synchronize()
print("20 steps after the device-side init: temperature", sim.compute_temperature())
