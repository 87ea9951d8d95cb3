"""Cost of recording the hit lists: the searching kernel, the recording kernel (forced at every evaluation) and the list walker.
    python tools/build_cost.py [workload] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sim = orbc.Simulation(bench.load_state(workload), kBT=0.22)
sim.run_langevin(4)
for name, opts, force in (("search", dict(nl_reuse=0), False), ("record (forced)", dict(nl_reuse=1), True), ("walk", dict(nl_reuse=1), False)):
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
    sim.profile_kernels(True)
    for _ in range(reps):
        if force:
            sim.set_option("nl_skin", 0.1)       # invalidates the lists: the next evaluation records
        sim.compute_pairwise_fused()
    rep = sim.kernel_report()
    sim.profile_kernels(False)
    print(name, sim.dump("nl_stats").tolist())
    for kn, n, us in rep[:8]:
        print(f"   {kn:40s} {n:4d} launches {us / n:8.1f} us mean")
