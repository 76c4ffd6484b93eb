#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share, count, average per kernel."""
import collections
import csv
import re
import sys


def main(path, per_step_div=1.0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        key = f"{name} grid={row.get('Grid Size', '')} blk={row.get('Block Size', '')}"
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}%  n={v[0]:4d}  avg={v[1] / v[0]:10.1f} us  {k[:130]}")


if __name__ == "__main__":
    main(sys.argv[1])
