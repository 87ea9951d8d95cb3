"""Print the SASS of one kernel from an ncu report with per-instruction executed counts and stall samples.
    python tools/ncu_sass.py report.ncu-rep kernel_regex [min_share]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + pat],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
start = [i for i, ln in enumerate(out) if ln.startswith('"Address"')][0]
rows = [r for r in csv.DictReader(out[start:]) if (r.get("Instructions Executed") or "").isdigit()]
tot_i = sum(int(r["Instructions Executed"]) for r in rows)
tot_s = sum(int(r["# Samples"]) for r in rows)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for k, r in enumerate(rows):
    ie, sm = int(r["Instructions Executed"]), int(r["# Samples"])
    if ie / tot_i < min_share and sm / max(tot_s, 1) < min_share:
        continue
    print(f"{k:4d} {ie / tot_i * 100:6.2f}% inst {sm / max(tot_s, 1) * 100:6.2f}% smp  thr {r['Avg. Threads Executed']:>5s}  {r['Source'].strip()}")
