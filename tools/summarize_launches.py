"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    agg[name][0] += ns
    agg[name][1] += 1
total = sum(v[0] for v in agg.values())
print(f"total {total/1e6:.3f} ms over {sum(v[1] for v in agg.values())} launches")
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:60]:60s} {n:8d} {ns/1e6:10.3f} {ns/total*100:6.1f}% {ns/n/1e3:9.1f}")
