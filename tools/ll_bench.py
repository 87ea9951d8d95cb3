"""The lipid pair kernels on a workload state: same hits (forces equal up to summation order), timings from CUDA event pairs.
    python tools/ll_bench.py [workload] [reps]
Variants: thread per lipid over candidate runs (search at every evaluation), the same with hit lists between rebuilds, and the
warp-per-cell tile kernel."""
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
sim.run_langevin(4)
base = None
for name, opts in (("run-list, search", dict(ll_variant=1, nl_reuse=0)), ("run-list, walking hit lists", dict(ll_variant=1, nl_reuse=1)),
                   ("run-list, gated, recording", dict(debug_nl_mode=1)), ("run-list, gated, searching", dict(debug_nl_mode=2)), ("tile", dict(debug_nl_mode=-1, ll_variant=0))):
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
    f = sim.download(0, "ft")
    if base is None:
        base = f
    den = np.linalg.norm(base["f"], axis=1) + np.sqrt((base["f"].astype(np.float64) ** 2).sum(1).mean())
    err = float((np.linalg.norm(f["f"].astype(np.float64) - base["f"], axis=1) / den).max())
    sim.profile_enable(True)
    for _ in range(reps):
        sim.compute_pairwise_fused()
    ms, n = sim.profile_read("pair_lipid")
    msp, npr = sim.profile_read("pair_protein")
    sim.profile_enable(False)
    print(f"{name:30s}: pair_lipid {ms / n * 1e3:7.1f} us  pair_protein {msp / max(npr, 1) * 1e3:7.1f} us; max rel err vs the first {err:.2e}", flush=True)
for name, opts in (("search", dict(ll_variant=1, nl_reuse=0)), ("hit lists", dict(ll_variant=1, nl_reuse=1))):
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.run_langevin(4); sim.synchronize()
    sim.event_record(0); sim.run_langevin(48); sim.event_record(1); sim.synchronize()
    print(f"{name:30s}: run_langevin {sim.event_elapsed_ms(0, 1) / 48 * 1e3:.1f} us/step   nl_stats {sim.dump('nl_stats').tolist()}", flush=True)
    sim.profile_kernels(True)
    sim.run_langevin(8)
    rep = sim.kernel_report()
    sim.profile_kernels(False)
    print(f"   kernel time {sum(r[2] for r in rep) / 8:.0f} us/step")
    for kn, n, us in rep[:14]:
        print(f"   {kn:40s} {n:4d} launches {us / n:8.1f} us mean")
