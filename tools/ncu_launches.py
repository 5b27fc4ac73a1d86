#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.

    tools/ncu_launches.py launches.csv "command line that was profiled" """
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    t = float(r[14].replace(",", "")) / 1e3
    a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
    a[0] += 1; a[1] += t
total = sum(a[1] for a in agg.values())
print("ncu launch list of: %s   (B200, round 1)" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
print("per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes")
print("%-60s %8s %12s %8s   %s" % ("kernel", "launches", "total us", "share", "block / grid"))
for name, (n, t, blk, grd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.1f %7.1f%%   %s / %s" % (name, n, t, 100 * t / total, blk, grd))
