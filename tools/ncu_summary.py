"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((int(r["ID"]), name, ns))
agg = defaultdict(lambda: [0, 0.0])
for _, n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns / 1e6:10.3f} ms {100 * ns / tot:5.1f}%  x{c:<6d} avg {ns / c / 1e3:9.2f} us  {n}")
