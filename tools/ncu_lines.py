"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py report.ncu-rep kernel_regex [min_share]"""
import collections
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
agg = collections.OrderedDict()
fname, hdr = None, None
for row in csv.reader(out):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) < len(hdr) or not row[0].isdigit():
        continue
    d = dict(zip(hdr, row))
    ie, sm = d.get("Instructions Executed", ""), d.get("# Samples", "")
    if not ie.isdigit():
        continue
    key = (fname, int(row[0]))
    a = agg.setdefault(key, [0, 0, row[1].strip()])
    a[0] += int(ie); a[1] += int(sm) if sm.isdigit() else 0
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp instructions {ti}, samples {ts}")
for (f, ln), a in sorted(agg.items()):
    if a[0] / ti >= min_share or a[1] / max(ts, 1) >= min_share:
        print(f"{f}:{ln:<5d} inst {a[0] / ti * 100:6.2f}%  smp {a[1] / max(ts, 1) * 100:6.2f}%  {a[2][:110]}")
