"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000.0 if unit in ("nsecond", "ns") else v * 1000.0 if unit in ("msecond", "ms") else v
    k = (name, row.get("Grid Size", ""))
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'total us':>10} {'n':>5} {'avg us':>9} {'share':>6}  kernel grid")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]:10.1f} {v[0]:5d} {v[1] / v[0]:9.1f} {100 * v[1] / tot:5.1f}%  {k[0][:70]} {k[1]}")
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
