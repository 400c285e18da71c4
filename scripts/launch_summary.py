"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/launch_summary.py launches.csv [divide_by_iterations] [top_n]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[row["Metric Unit"]]
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
own = sum(v[1] for k, v in agg.items() if "wf::" in k)
print(f"# {path}: {n // iters} launches, {tot / iters:.1f} us per forward (serialised, cold-cache ncu times); "
      f"own kernels (wf::) {100 * own / tot:.1f}% of the time")
for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t / iters:9.1f} us {100 * t / tot:5.1f}% x{c // iters:4d}  {name}")
