"""Per-call device timings on a workload state (CUDA events on the context's stream).
    python tools/kernel_bench.py [workload] [reps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
st = bench.load_state(workload)
sim = orbc.Simulation(st, kBT=0.22)
sim.set_option("nl_reuse", 0)       # per-call timings of the searching kernels (tools/ll_bench.py times the hit lists)
sim.run_langevin(4)


def timeit(name, fn, reps=reps):
    fn(); sim.synchronize()
    sim.event_record(0)
    for _ in range(reps):
        fn()
    sim.event_record(1); sim.synchronize()
    print(f"{name:34s} {sim.event_elapsed_ms(0, 1) / reps * 1e3:10.1f} us", flush=True)


ref = None
for impl in (1, 2):
    sim.set_option("pair_impl", impl)
    sim.clear_force(); sim.compute_pairwise_fused()
    f = [sim.download(s, "ft") for s in (0, 1)]
    if ref is None:
        ref = f
    else:
        for s in (0, 1):
            for k in "ft":
                a, b = f[s][k].astype(np.float64), ref[s][k].astype(np.float64)
                den = np.linalg.norm(b, axis=1) + np.sqrt((b * b).sum(1).mean()) + 1e-30
                print(f"  impl {impl} vs 1: species {s} {k} max rel {np.max(np.linalg.norm(a - b, axis=1) / den) if len(a) else 0:.2e}")
    timeit(f"compute_pairwise_fused impl={impl}", sim.compute_pairwise_fused)
    sim.profile_enable(True)
    for _ in range(reps):
        sim.compute_pairwise_fused()
    for k in ("pair_lipid", "pair_protein"):
        ms, n = sim.profile_read(k)
        print(f"    {k:30s} {ms / max(n, 1) * 1e3:10.1f} us")
    sim.profile_enable(False)
sim.clear_force()
timeit("compute_bonded", sim.compute_bonded)
sim.clear_force()
timeit("verlet_langevin", sim.verlet_langevin)
sim.nstep = 2
timeit("rebuild (no Morton sort)", sim.rebuild)
sim.nstep = 24
timeit("rebuild (Morton sort)", sim.rebuild)
sim.nstep = 2
timeit("voronoi_update", sim.voronoi_update)
timeit("cell_update lipid", lambda: sim.cell_update(0))
timeit("cell_update protein", lambda: sim.cell_update(1))
timeit("compute_temperature", sim.compute_temperature)
sim.nstep = 0
timeit("run_langevin(2)", lambda: sim.run_langevin(2))
sim.set_option("pair_impl", 2)
for lanes in (1, 2, 4):
    sim.set_option("prot_lanes", lanes)
    for frac in (1.0, 0.125):
        sim.set_option("debug_owned_fraction", frac)
        sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
        sim.profile_enable(True)
        for _ in range(reps):
            sim.compute_pairwise_fused()
        print(f"prot_lanes {lanes}, owned fraction {frac}: pair_protein {sim.profile_read('pair_protein')[0] / reps * 1e3:.1f} us")
        sim.profile_enable(False)
sim.set_option("debug_owned_fraction", 1.0)
sim.set_option("prot_lanes", 0)
# how the pair kernels scale with the number of owned particles (what one rank of a decomposed run sees)
sim.set_option("pair_impl", 2)
for frac in (1.0, 0.5, 0.25, 0.125):
    sim.set_option("debug_owned_fraction", frac)
    sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
    sim.profile_enable(True)
    for _ in range(reps):
        sim.compute_pairwise_fused()
    print(f"owned fraction {frac}: " + "  ".join(f"{k} {sim.profile_read(k)[0] / reps * 1e3:.1f} us" for k in ("pair_lipid", "pair_protein")))
    sim.profile_enable(False)
sim.set_option("debug_owned_fraction", 1.0)
