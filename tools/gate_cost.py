"""What the three-way gated launch costs: per-kernel times (event pairs around every launch) of force evaluations in a fixed mode.
    python tools/gate_cost.py [workload] [option=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

st = bench.load_state(sys.argv[1] if len(sys.argv) > 1 else "rbc")
sim = orbc.Simulation(st, kBT=0.22)
for kv in sys.argv[2:]:
    sim.set_option(kv.split("=")[0], float(kv.split("=")[1]))
sim.run_langevin(4)
for name, opts in (("no lists", dict(nl_reuse=0)), ("walking", dict(nl_reuse=1)), ("recording", dict(debug_nl_mode=1)), ("searching", dict(debug_nl_mode=2))):
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
    sim.profile_kernels(True)
    for _ in range(8):
        sim.compute_pairwise_fused()
    rep = sim.kernel_report()
    sim.profile_kernels(False)
    print(f"{name}: kernel time {sum(r[2] for r in rep) / 8:.0f} us per evaluation")
    for kn, n, us in rep:
        print(f"   {kn:40s} {n:4d} launches {us / n:8.1f} us mean")
