#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) launches, total ms, share.

    python profiles/tools/launch_list_summary.py gpurun_out/X_launches.csv > profiles/r2/X_launches.summary.txt
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
K, G, MN, MV, MU = (hdr.index(c) for c in ("Kernel Name", "Grid Size", "Metric Name", "Metric Value", "Metric Unit"))
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= MV or r[MN] != "gpu__time_duration.sum":
        continue
    v = float(r[MV].replace(",", ""))
    u = r[MU]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
    key = (r[K].split("(")[0].replace("dmpc::", ""), r[G])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("kernel | grid | launches | total ms | share")
for (k, g), (n, ms) in agg.items():
    print("%s | %s | %d | %.3f | %.1f%%" % (k, g, n, ms, 100.0 * ms / tot))
