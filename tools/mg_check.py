"""Decomposed run driven from ONE process (one host thread per rank; ranks are spread over the visible GPUs, several ranks
may share a device): checks it against the single-GPU run on the same state and prints device timings.
    python tools/mg_check.py [workload] [world] [steps]
"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402
import torch  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ndev = torch.cuda.device_count()
st = bench.load_state(workload)
print(f"{workload}: {len(st['lx'])} lipids, {len(st['px'])} proteins, {len(st['centroids'])} cells; {world} ranks on {ndev} device(s)", flush=True)

one = orbc.Simulation(st, kBT=0.22)
if len(st["px"]) // world <= 300000:
    one.set_option("prot_lanes", 4)       # same lanes per protein (= same summation order) as the ranks will choose
one.run_langevin(steps)
one.synchronize()
ref = [one.download(s, "xvno") for s in (0, 1)]
ref_cs = [one.dump("cell_start_l"), one.dump("cell_start_p")]
one.close()

sims = [orbc.Simulation(st, kBT=0.22, rank=r, world=world, device=r % ndev) for r in range(world)]
blobs = [s.mg_export() for s in sims]
for s in sims:
    s.mg_connect(blobs)
ms = [0.0] * world


def on_all(fn):
    errs = []

    def work(s):
        try:
            fn(s)
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(s,)) for s in sims]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]


def timed(s, n):
    s.event_record(0)
    s.run_langevin(n)
    s.event_record(1)
    s.synchronize()
    ms[s.rank] = s.event_elapsed_ms(0, 1)


on_all(lambda s: timed(s, steps))
worst = 0.0
for sp in (0, 1):
    parts = {f: [] for f in "xvno"}
    for s in sims:
        b, e = s.owned_range(sp)
        d = s.download(sp, "xvno")
        for f in "xvno":
            parts[f].append(d[f][b:e])
    for f in "xvno":
        got = np.concatenate(parts[f])
        assert got.shape == ref[sp][f].shape, (got.shape, ref[sp][f].shape)
        err = float(np.abs(got - ref[sp][f]).max()) if len(got) else 0.0
        worst = max(worst, err / (1 + float(np.abs(ref[sp][f]).max()) if len(got) else 1))
        print(f"  species {sp} {f}: max |diff| {err:.3e}")
for s in sims:
    assert (s.dump("cell_start_l") == ref_cs[0]).all() and (s.dump("cell_start_p") == ref_cs[1]).all(), f"cell_start differs on rank {s.rank}"
print(f"cell_start identical on all ranks; worst relative deviation {worst:.2e}")
print("first run (cold):", " ".join(f"{m / steps * 1e3:.0f}" for m in ms), "us/step per rank")
for rep in range(2):
    t0 = time.perf_counter()
    on_all(lambda s: timed(s, 40))
    wall = time.perf_counter() - t0
    n = len(st["lx"]) + len(st["px"])
    print(f"40 steps: max over ranks {max(ms) / 40 * 1e3:.0f} us/step -> {n * 40 / (max(ms) * 1e-3) / 1e9:.3f} G particle-steps/s (wall {wall * 1e3 / 40:.3f} ms/step)", flush=True)
def barriers(s, n):
    s.event_record(0)
    s.set_option("debug_barriers", n)
    s.event_record(1)
    s.synchronize()
    ms[s.rank] = s.event_elapsed_ms(0, 1)


on_all(lambda s: barriers(s, 200))
on_all(lambda s: barriers(s, 200))
print(f"200 back-to-back barriers: {max(ms) / 200 * 1e3:.2f} us each")
sims[0].profile_enable(True)
on_all(lambda s: timed(s, 20))
print("rank 0 per class (us/step):", {k: round(sims[0].profile_read(k)[0] / 20 * 1e3, 1) for k in orbc.engine.PROF})
for s in sims:
    s.profile_kernels(True)
on_all(lambda s: timed(s, 24))
for s in sims[:1] + sims[-1:]:
    rep = s.kernel_report()
    tot = sum(r[2] for r in rep)
    print(f"rank {s.rank}: kernel time {tot / 24:.0f} us/step of {ms[s.rank] / 24 * 1e3:.0f} us/step elapsed")
    for name, n, us in rep[:22]:
        print(f"   {name:24s} {n:5d} launches {us / n:8.1f} us mean {us / 24:8.1f} us/step")
for s in sims:
    s.close()
