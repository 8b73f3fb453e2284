#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (and by kernel + grid)."""
import collections
import csv
import re
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    per = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, vi, gi, ui = r.index("Kernel Name"), r.index("Metric Value"), r.index("Grid Size"), r.index("Metric Unit")
            continue
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name).replace("dc::", "").replace("void ", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1000.0
        elif r[ui] in ("ms", "msecond"):
            v *= 1000.0
        agg[name][0] += 1
        agg[name][1] += v
        per[(name, r[gi])][0] += 1
        per[(name, r[gi])][1] += v
    tot = sum(v[1] for v in agg.values())
    print("total us %.1f over %d launches" % (tot, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-40s n=%5d us=%9.1f avg=%7.1f" % (k, v[0], v[1], v[1] / v[0]))
    print()
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-34s %-20s n=%5d us=%9.1f avg=%7.1f" % (k[0], k[1], v[0], v[1], v[1] / v[0]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
