"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share).
    python tools/ncu_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def summarise(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


if __name__ == "__main__":
    agg = summarise(sys.argv[1])
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.3f} |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
