"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (and grid): count, total us, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
by_grid = len(sys.argv) > 2
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = d["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    name = d["Kernel Name"].split("(")[0][:48]
    if by_grid:
        name += " grid=" + d["Grid Size"]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% {v[0]:5d}  {k}")
print(f"{tot:10.1f} us total, {sum(v[0] for v in agg.values())} launches")
