"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
agg = defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)[:90]
    agg[short][0] += 1
    agg[short][1] += ns
total = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {total / 1e6:.3f} ms (serialised, cold cache: compare SHARES)")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{ns / 1e6:9.3f} ms  {100 * ns / total:5.1f}%  x{n:<5d} {name}")
