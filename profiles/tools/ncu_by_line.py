#!/usr/bin/env python
"""Per-CUDA-source-line share of the warp-stall samples of one kernel.

    ncu -i X.ncu-rep --page source --csv > sass.csv          # per-SASS-instruction samples
    cuobjdump -xelf all build/lqr_launch_f64.o && nvdisasm -gi lqr_launch.sm_100a.cubin > lines.txt
    python profiles/tools/ncu_by_line.py sass.csv lines.txt <mangled kernel name> <source file>

Joins the two listings by instruction order (same binary) and attributes every instruction to the outermost
line of <source file> on its inline stack."""
import collections
import csv
import re
import sys

sass_csv, lines_txt, kernel, srcfile = sys.argv[1:5]
insts, cur, active = [], None, False
for ln in open(lines_txt):
    if ln.startswith(".text."):
        active = ln.strip().rstrip(":") == ".text." + kernel
        continue
    if not active:
        continue
    if "//## File" in ln:
        locs = re.findall(r'"([^"]+)", line (\d+)', ln)
        mine = [(f, int(l)) for f, l in locs if f.endswith(srcfile)]
        cur = mine[-1] if mine else (locs[-1][0], int(locs[-1][1]))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        insts.append((cur, m.group(2)))
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
N = hdr.index("# Samples")
assert len(insts) == len(data), (len(insts), len(data))
agg = collections.Counter()
for (loc, _), r in zip(insts, data):
    agg[loc] += int(r[N])
tot = sum(agg.values())
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    text = ""
    if k:
        try:
            src.setdefault(k[0], open(k[0]).read().splitlines())
            text = src[k[0]][k[1] - 1].strip()[:100]
        except Exception:
            pass
    print("%5.1f%%  %s:%s  %s" % (100.0 * v / tot, k[0].split("/")[-1] if k else None, k[1] if k else None, text))
