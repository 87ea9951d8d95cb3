"""Small driver for compute-sanitizer (tools/sanitize.sh): the whole hot path on a golden fixture, single GPU and decomposed over
ranks that share the device.    python tools/sanitize_driver.py [world] [steps]"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openrbc_b200 as orbc  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
opts = dict(kv.split("=") for kv in sys.argv[3:])
dt = float(opts.pop("dt", 1e-3))
ndev = int(opts.pop("ndev", 1))
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "vesicle_ico0.npz")))
st = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
if world == 1:
    sim = orbc.Simulation(st, dt=dt, kBT=0.22)
    for k, v in opts.items():
        sim.set_option(k, float(v))
    sim.run_langevin(steps); sim.run_nh(2); sim.run_minimize(2)
    sim.delete_lipid(2.5); sim.compute_temperature(); sim.constrain_volume(3.15, 0.05); sim.save_frame(31)
    for v in (0, 1):
        sim.set_option("ll_variant", v); sim.compute_pairwise_fused()
    sim.synchronize()
    print("single ok", sim.dump("nl_stats").tolist())
else:
    sims = [orbc.Simulation(st, dt=dt, kBT=0.22, rank=r, world=world, device=r % ndev) for r in range(world)]
    for s in sims:
        for k, v in opts.items():
            s.set_option(k, float(v))
    blobs = [s.mg_export() for s in sims]
    for s in sims:
        s.mg_connect(blobs)
    errs = []

    def work(s):
        try:
            s.run_langevin(steps); s.run_nh(2); s.run_minimize(2); s.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(s,)) for s in sims]
    [t.start() for t in th]; [t.join() for t in th]
    if errs:
        raise errs[0]
    print("decomposed ok", [s.dump("nl_stats").tolist() for s in sims])
