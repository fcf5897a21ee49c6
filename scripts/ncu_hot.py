"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (stdin).
usage: ncu -i rep --page source --csv | python scripts/ncu_hot.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[hi + 1:]):
    try:
        data.append((int(r[isamp]), k, r[isrc].strip(), int(r[iex])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(f"{len(data)} SASS instructions, {tot} samples")
for s, k, src, ex in sorted(data, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  #{k:5d}  exec={ex:9d}  {src[:110]}")
