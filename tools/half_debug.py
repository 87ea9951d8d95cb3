"""Diagnose k_pair_ll_h against k_pair_ll on a workload: which lipids differ, and which partner the prefilter treats differently."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openrbc_b200 as orbc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rbc"
sim = orbc.Simulation(bench.load_state(workload), kBT=0.22)
sim.run_langevin(4)
out = {}
for half in (0, 1):
    sim.set_option("ll_half", half)
    sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
    out[half] = sim.download(0, "ft")
d = np.abs(out[1]["f"] - out[0]["f"]).max(axis=1)
bad = np.nonzero(d > 0)[0]
print(f"{len(bad)} of {len(d)} lipids differ; max |df| {d.max():.3e}; |f| max {np.abs(out[0]['f']).max():.3e}")
if len(bad) == 0:
    sys.exit(0)
print("first differing slots:", bad[:20], "parity histogram:", np.bincount(bad & 1, minlength=2))
x = sim.download(0, "x", affiliation=True)
cell = x["affiliation"]; X = x["x"].astype(np.float32)
cs = sim.dump("cell_start_l"); cen = sim.dump("centroids"); cnt = sim.dump("stencil_counts"); st = sim.dump("stencil")
for i in bad[:6]:
    c = cell[i]
    cand = np.concatenate([np.arange(cs[c2], cs[c2 + 1]) for c2 in st[c, :cnt[c, 0]]])
    c2s = np.concatenate([np.full(cs[c2 + 1] - cs[c2], c2) for c2 in st[c, :cnt[c, 0]]])
    dd = X[i] - X[cand]
    r2 = (dd * dd).sum(1)
    hit = (r2 < np.float32(6.76)) & (r2 > 1e-5)
    # emulate the prefilter
    o = cen[c2s]
    xh = (X[i] - o).astype(np.float16)
    q = (X[cand] - o).astype(np.float16)
    dh = (xh - q).astype(np.float16)
    r2h = (dh[:, 0] * dh[:, 0]).astype(np.float16)
    r2h = (dh[:, 1].astype(np.float32) * dh[:, 1].astype(np.float32) + r2h.astype(np.float32)).astype(np.float16)
    r2h = (dh[:, 2].astype(np.float32) * dh[:, 2].astype(np.float32) + r2h.astype(np.float32)).astype(np.float16)
    pre = r2h < np.float16(6.96)
    print(f"slot {i} cell {c} ({cs[c]}..{cs[c+1]}): {hit.sum()} hits of {len(cand)} candidates; prefilter passes {pre.sum()}, misses {(hit & ~pre).sum()} hits; "
          f"max |x - o| {np.abs(X[i] - o).max():.2f}, max |rel partner| {np.abs(X[cand] - o).max():.2f}; df {out[1]['f'][i] - out[0]['f'][i]}")
    miss = np.nonzero(hit & ~pre)[0]
    for m in miss[:3]:
        print("    missed partner slot", cand[m], "cell", c2s[m], "r2", r2[m], "r2h", r2h[m])
