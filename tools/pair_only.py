"""Minimal driver for ncu captures: load a state, advance a few steps, then N force evaluations.
    python tools/pair_only.py [workload] [n] [pair_impl] [ll_variant] [debug_nl_mode: 1 every evaluation records the hit lists, 2 searches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sim = orbc.Simulation(bench.load_state(workload), kBT=0.22)
if len(sys.argv) > 3:
    sim.set_option("pair_impl", int(sys.argv[3]))
if len(sys.argv) > 4:
    sim.set_option("ll_variant", int(sys.argv[4]))
sim.run_langevin(4)
if len(sys.argv) > 5:
    sim.set_option("debug_nl_mode", int(sys.argv[5]))
for _ in range(n):
    sim.compute_pairwise_fused()
sim.synchronize()
