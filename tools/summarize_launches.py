#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total/avg us, share."""
import csv, re, sys
from collections import defaultdict

rows = defaultdict(lambda: [0, 0.0])
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void |\(anonymous namespace\)::", "", name)
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    rows[name][0] += 1
    rows[name][1] += us
tot = sum(v[1] for v in rows.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.0f | %.1f | %.3f |" % (k, c, t, t / c, t / tot))
print("\nTotal %.1f ms over %d launches" % (tot / 1e3, sum(v[0] for v in rows.values())))
