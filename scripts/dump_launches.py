"""Per-launch list of the last N launches of an ncu gpu__time_duration CSV: index, kernel, grid, block, microseconds.
usage: dump_launches.py launches.csv last_n"""
import csv
import re
import sys

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
rows = rows[-int(sys.argv[2]):]
for i, r in enumerate(rows):
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    print(f"{i:4d} {re.sub(r'[(].*', '', r['Kernel Name'])[:44]:44s} grid {r.get('Grid Size', '?'):>18s} block {r.get('Block Size', '?'):>14s} {us:10.1f} us")
