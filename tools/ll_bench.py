"""Variants of the lipid pair kernel on a workload state: same hits (forces equal up to summation order), timings from CUDA event pairs.
    python tools/ll_bench.py [workload] [reps] [variants...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = [int(v) for v in sys.argv[3:]] or [1, 0]     # 1: thread-per-lipid run-list kernel, 0: warp-per-cell tile kernel
st = bench.load_state(workload)
sim = orbc.Simulation(st, kBT=0.22)
sim.run_langevin(4)
base = None
for var in variants:
    sim.set_option("ll_variant", var)
    sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
    f = sim.download(0, "ft")
    if base is None:
        base = f
    nl = len(f["f"])
    # protein -> lipid reactions arrive by atomics in any order: compare the lipids no protein touched bit for bit
    same_f = (f["f"] == base["f"]).all(axis=1); same_t = (f["t"] == base["t"]).all(axis=1)
    sim.profile_enable(True)
    for _ in range(reps):
        sim.compute_pairwise_fused()
    ms, n = sim.profile_read("pair_lipid")
    sim.profile_enable(False)
    import numpy as np
    den = np.linalg.norm(base["f"], axis=1) + np.sqrt((base["f"].astype(np.float64) ** 2).sum(1).mean())
    err = float((np.linalg.norm(f["f"].astype(np.float64) - base["f"], axis=1) / den).max())
    print(f"ll_variant {var}: pair_lipid {ms / n * 1e3:.1f} us; vs variant {variants[0]}: max rel err {err:.2e}, rows bit-identical f {same_f.mean():.6f} t {same_t.mean():.6f}", flush=True)
for var in variants:
    sim.set_option("ll_variant", var)
    sim.run_langevin(4); sim.synchronize()
    sim.event_record(0); sim.run_langevin(48); sim.event_record(1); sim.synchronize()
    print(f"ll_variant {var}: run_langevin {sim.event_elapsed_ms(0, 1) / 48 * 1e3:.1f} us/step", flush=True)
