#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean us."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][:64]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':64s} {'n':>5s} {'mean_us':>10s} {'min_us':>9s} {'max_us':>9s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:64s} {len(v):5d} {sum(v)/len(v):10.2f} {min(v):9.2f} {max(v):9.2f} {100*sum(v)/tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
