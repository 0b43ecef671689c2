#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: python profiles/aggregate_launches.py gpurun_out/X.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("%-44s %7s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %7d %12.1f %10.2f %6.1f%%" % (k[:44], v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
