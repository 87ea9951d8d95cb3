"""The end-to-end leg of bench.py (per-call loop + cleanup every 60 steps + temperature every 100) on a decomposed run driven from
ONE process (one host thread per rank; ranks may share a device); prints every call that takes unusually long.
    python tools/e2e_mg_check.py [workload] [world] [steps]
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402
import torch  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 240
ndev = torch.cuda.device_count()
st = bench.load_state(workload)
sims = [orbc.Simulation(st, kBT=0.22, rank=r, world=world, device=r % ndev) for r in range(world)]
blobs = [s.mg_export() for s in sims]
for s in sims:
    s.mg_connect(blobs)
lock = threading.Lock()


def loop(s):
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        what = "step"
        if s.nstep % 60 == 0:
            n = s.delete_lipid(2.5); what = f"delete_lipid -> {n} + step"
        s.step_langevin_checked()
        if s.nstep % 100 == 0:
            s.compute_temperature(); what += " + temperature"
        ms = (time.perf_counter() - t0) * 1e3
        if ms > 20.0 or what != "step":
            with lock:
                print(f"rank {s.rank} nstep {s.nstep - 1}: {what} {ms:.2f} ms", flush=True)
    with lock:
        print(f"rank {s.rank}: {steps} steps in {(time.perf_counter() - t_all) * 1e3:.1f} ms, {s.size(0)} lipids", flush=True)


th = [threading.Thread(target=loop, args=(s,)) for s in sims]
[t.start() for t in th]
[t.join() for t in th]
