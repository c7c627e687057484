"""Histogram of executed SASS instructions / stall samples from `ncu --page source --csv`."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[1]
ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, samp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    if len(r) <= ie or not r[ie].strip():
        continue
    try:
        n, s = int(float(r[ie])), int(float(r[isamp] or 0))
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
    op = ".".join((m.group(2) if m else r[ia][:10]).split(".")[:2])
    ops[op] += n
    samp[op] += s
    tot += n
    stot += s
print("total warp instr", tot, "per tile", tot / tiles)
for op, n in ops.most_common(45):
    print(f"{op:24s} {n:11d} {n / tot * 100:5.1f}%  per-tile {n / tiles:8.0f}  samples {samp[op] / max(stot,1) * 100:5.1f}%")
