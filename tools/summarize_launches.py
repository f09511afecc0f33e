"""Summarise an ncu --csv launch list (gpu__time_duration.sum + dram bytes per kernel launch) into per-kernel averages.
Usage: python tools/summarize_launches.py <raw.csv>  -> JSON on stdout."""
import csv
import json
import re
import sys
from collections import defaultdict

rows = defaultdict(dict)
names = {}
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)   # -> us
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    rows[r["ID"]][m] = v
    k = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1]
    names[r["ID"]] = k
agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "dram": 0.0})
for i, m in rows.items():
    a = agg[names[i]]
    a["launches"] += 1
    a["us"] += m.get("gpu__time_duration.sum", 0.0)
    a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
total = sum(a["us"] for a in agg.values()) or 1.0
out = {k: {"launches": a["launches"], "avg_us": a["us"] / a["launches"], "share": a["us"] / total,
           "dram_mb_per_launch": a["dram"] / a["launches"] / 1e6}
       for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])}
json.dump(out, sys.stdout, indent=1)
