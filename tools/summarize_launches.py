"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv): per-kernel launches, total time,
share, DRAM bytes.  usage: python tools/summarize_launches.py launches.csv [--json out.json --skip-first-pass]"""
import csv
import json
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    rows.append(r)
per = defaultdict(lambda: defaultdict(float))
ids = {}
for r in rows:
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        per[name]["ms"] += val * scale
        per[name]["n"] += 1
    elif m.startswith("dram__bytes"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[name]["dram"] += val * scale
tot = sum(v["ms"] for v in per.values())
print(f"{path}: {int(sum(v['n'] for v in per.values()))} launches, {tot:.2f} ms of kernel time")
for name, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {v['ms']:10.3f} ms  {100 * v['ms'] / tot:5.1f} %  {int(v['n']):5d} launches  {v['dram'] / 1e6:10.1f} MB DRAM   {name[:110]}")
if "--json" in sys.argv:
    out = sys.argv[sys.argv.index("--json") + 1]
    json.dump({k: dict(v) for k, v in per.items()}, open(out, "w"), indent=1)
