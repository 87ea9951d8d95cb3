"""Rebuild timings on a workload state (CUDA events) and the per-kernel list of two rebuilds.   python tools/rebuild_bench.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, openrbc_b200 as orbc
sim = orbc.Simulation(bench.load_state(sys.argv[1] if len(sys.argv) > 1 else "rbc"), kBT=0.22)
sim.run_langevin(4)
for nstep, what in ((2, "no Morton sort"), (24, "Morton sort")):
    sim.nstep = nstep
    sim.rebuild(); sim.synchronize()
    sim.event_record(0)
    for _ in range(10):
        sim.rebuild()
    sim.event_record(1); sim.synchronize()
    print(f"rebuild ({what}): {sim.event_elapsed_ms(0, 1) / 10 * 1e3:.1f} us", flush=True)
sim.nstep = 2
sim.profile_kernels(True)
sim.rebuild(); sim.rebuild()
for name, n, us in sim.kernel_report():
    print(f"   {name:28s} {n:3d} launches {us / n:8.1f} us mean")
sim.profile_kernels(False)
sim.nstep = 4
sim.run_langevin(4); sim.synchronize()
sim.event_record(0); sim.run_langevin(48); sim.event_record(1); sim.synchronize()
print(f"run_langevin {sim.event_elapsed_ms(0, 1) / 48 * 1e3:.1f} us/step", flush=True)
