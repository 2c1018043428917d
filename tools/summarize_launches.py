"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list: time share per kernel and, when the DRAM counters were collected, measured DRAM bytes per launch.

    python tools/summarize_launches.py launches.csv [traffic.json]

traffic.json (optional) receives {"gemm_kernel": {"launches", "dram_bytes_per_launch", ...}} -- the file
bench.py reads for `roofline.traffic` (profiles/gemm_traffic.json)."""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT_NS = {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}
UNIT_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
per_id = defaultdict(dict)  # launch id -> {name, ns, rd, wr}
for r in csv.DictReader(lines):
    d = per_id[r["ID"]]
    d["name"] = r["Kernel Name"]
    m, v, u = r.get("Metric Name"), float(r["Metric Value"].replace(",", "")), r.get("Metric Unit", "")
    if m == "gpu__time_duration.sum":
        d["ns"] = v * UNIT_NS.get(u, 1)
    elif m == "dram__bytes_read.sum":
        d["rd"] = v * UNIT_B.get(u, 1)
    elif m == "dram__bytes_write.sum":
        d["wr"] = v * UNIT_B.get(u, 1)
rows = [d for d in per_id.values() if "ns" in d]
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in rows:
    short = re.sub(r"\(.*", "", d["name"])
    short = re.sub(r"^void ", "", short)[:90]
    a = agg[short]
    a[0] += 1
    a[1] += d["ns"]
    a[2] += d.get("rd", 0.0)
    a[3] += d.get("wr", 0.0)
total = sum(v[1] for v in agg.values())
have_dram = any("rd" in d for d in rows)
print(f"launches {len(rows)}  total {total / 1e6:.3f} ms (serialised, cold cache: compare SHARES)")
for name, (n, ns, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    extra = f"  dram r {rd / 1e6:9.1f} MB  w {wr / 1e6:9.1f} MB  -> {(rd + wr) / ns:6.2f} GB/ms" if have_dram else ""
    print(f"{ns / 1e6:9.3f} ms  {100 * ns / total:5.1f}%  x{n:<5d} {name}{extra}")
if have_dram and len(sys.argv) > 2:
    g = [v for k, v in agg.items() if "gemm_kernel" in k]
    n = sum(v[0] for v in g)
    out = {"gemm_kernel": {"launches": n, "dram_bytes_per_launch": sum(v[2] + v[3] for v in g) / max(n, 1),
                           "dram_read_bytes": sum(v[2] for v in g), "dram_write_bytes": sum(v[3] for v in g),
                           "ncu_ms_total": sum(v[1] for v in g) / 1e6,
                           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                                     "--clock-control none over one steady-state pre-training step "
                                     "(bench.py --profile-step), all gemm_kernel launches"},
           "all_kernels": {"launches": len(rows), "dram_bytes": sum(v[2] + v[3] for v in agg.values()),
                           "ncu_ms_total": total / 1e6}}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print("wrote", sys.argv[2])
