"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.
usage: summarize_launches.py launches.csv [last_n]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
if len(sys.argv) > 2:  # keep only the last N launches (e.g. the last k step of the timed pass)
    rows = rows[-int(sys.argv[2]):]
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1
    agg[k][1] += v
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (serialised, cold-cache: compare SHARES)")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {v:12.1f} {v/n:10.1f} {100*v/tot:6.1f}%")
