#!/usr/bin/env python
"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_summary.py gpurun_out/<tag>_launches.csv > profiles/<name>.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("ufo_b200::", "")
    ns = float(r[14].replace(",", ""))
    if r[13] == "us":
        ns *= 1e3
    elif r[13] == "ms":
        ns *= 1e6
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += ns
tot = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print("# every kernel launch of the run (cold-cache, serialised by ncu): compare SHARES, not absolutes")
print("%-34s %6s %10s %10s %7s" % ("kernel", "count", "total_ms", "mean_ms", "share"))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %6d %10.3f %10.4f %6.1f%%" % (name, n, ns / 1e6, ns / 1e6 / n, 100 * ns / tot))
