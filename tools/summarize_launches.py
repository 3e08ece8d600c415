"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms (serialised, cold-cache: compare shares)")
print(f"{'kernel':70s} {'n':>5s} {'ms':>10s} {'share':>7s}")
for n, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {cnt:5d} {ns / 1e6:10.3f} {100 * ns / tot:6.2f}%")
