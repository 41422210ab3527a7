"""ncu --csv launch list -> per-kernel count / total / mean time.  usage: summarise_launches.py file.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |ldu::|\(anonymous namespace\)::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}.get(unit, 1e-3)
    tot[name][0] += 1
    tot[name][1] += v
all_us = sum(v[1] for v in tot.values())
print(f"{sum(v[0] for v in tot.values())} launches, {all_us:.1f} us in kernels")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{us:10.1f} us {100*us/all_us:5.1f}%  n={n:5d}  mean {us/n:8.2f} us  {name[:110]}")
