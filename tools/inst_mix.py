#!/usr/bin/env python
"""Instruction mix of one kernel from an ncu report's source page (SASS):
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > k.csv; tools/inst_mix.py k.csv POINTS"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 23
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr, data = rows[hi], rows[hi + 1:]
ia, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops, st, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia].strip())
    if not m:
        continue
    op = m.group(2).split('.')[0]
    if op in ('LDS', 'STS', 'LDG', 'STG', 'LD', 'ST'):
        op = '.'.join(m.group(2).split('.')[:1] + [p for p in m.group(2).split('.')[1:] if p in ('64', '128', 'U8', 'S8')])
    n = int(r[ie]); ops[op] += n; tot += n; st[op] += int(r[iss] or 0)
print("total warp inst", tot, " thread-inst/pt %.1f" % (tot * 32 / pts), " stall samples", sum(st.values()))
for op, n in ops.most_common(30):
    print(f"{op:14s} {n:10d} {n*32/pts:6.2f}/pt  stall {st[op]}")
