"""BASELINE.json configs[4]: Nose-Hoover run (orbc_run_nh) with the periodic Morton reorder and a trajectory frame every 100 steps.
Three ways to get the frames out: none (pure compute), the device-assembled frame with the overlapped copy (orbc_save_frame_begin /
_end: what the drop-in program uses), and the per-array download a host-side save_frame needs (orbc_download).
    python tools/config5_nh_frames.py [workload] [steps] [every]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 100
st = bench.load_state(workload)
n = len(st["lx"]) + len(st["px"])
out = open(os.devnull, "wb")
import torch  # noqa: E402
pinned = torch.empty(64 * n + 4096, dtype=torch.uint8, pin_memory=True).numpy()      # caller-owned pinned buffer for the synchronous path
for mode in ("no frames", "device frame, overlapped copy", "device frame, synchronous", "per-array download"):
    sim = orbc.Simulation(st, kBT=0.22)
    sim.zeta = 0.0
    sim.run_nh(20)
    if mode.startswith("device frame, overlapped"):      # the two pinned buffers are allocated on first use: not part of the steady state
        sim.save_frame_begin(7); sim.save_frame_begin(7); sim.save_frame_end(); sim.save_frame_end()
    sim.synchronize()
    nbytes = 0
    t0 = time.perf_counter()
    sim.event_record(0)
    for k in range(steps // every):
        sim.run_nh(every)
        if mode == "device frame, overlapped copy":
            if k:
                fr = sim.save_frame_end(); out.write(memoryview(fr)); nbytes += fr.nbytes
            sim.save_frame_begin(7)
        elif mode == "device frame, synchronous":
            fr = sim.save_frame(7, out=pinned); out.write(memoryview(fr)); nbytes += fr.nbytes
        elif mode == "per-array download":
            for s in (0, 1):
                d = sim.download(s, "xn", ids=(s == 1), affiliation=True)
                nbytes += sum(v.nbytes for v in d.values())
    if mode == "device frame, overlapped copy":
        fr = sim.save_frame_end(); out.write(memoryview(fr)); nbytes += fr.nbytes
    sim.event_record(1); sim.synchronize()
    wall = time.perf_counter() - t0
    print(f"{mode:32s}: {wall / steps * 1e3:7.4f} ms/step wall ({sim.event_elapsed_ms(0, 1) / steps:7.4f} device), {n * steps / wall / 1e9:6.3f} G particle-steps/s, "
          f"{nbytes / 1e6:8.1f} MB of frames, zeta {sim.zeta:.5f}", flush=True)
    sim.close()
