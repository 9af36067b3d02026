#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = r[ki].split("(")[0]
    tot[name][0] += 1
    tot[name][1] += us
total = sum(t for _, t in tot.values())
print("%d launches, %.1f ms total device time (serialised under ncu: compare SHARES, not absolutes)" % (sum(c for c, _ in tot.values()), total / 1e3))
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%7d launches %11.1f us  %5.1f %%  %s" % (c, t, 100 * t / total, name))
# per grid size of the round kernels: which rounds the time goes to
if "Grid Size" in hdr:
    gi = hdr.index("Grid Size")
    by = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if "k_round" not in r[ki]:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        key = (r[ki].split("(")[0].replace("void ", ""), r[gi])
        by[key][0] += 1
        by[key][1] += us
    print("round kernels by grid size:")
    for (name, grid), (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:24]:
        print("%7d launches %11.1f us  %7.1f us each  grid %-14s %s" % (c, t, t / c, grid, name))
