#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: instruction counts by opcode and the hottest SASS lines.
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_top.py [N]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
ia, isrc, isamp, iex, ithr = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
body = rows[hdr_i + 1:]
ops, samp = Counter(), Counter()
tot = tots = 0
for r in body:
    op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    ex, sa = int(r[iex]), int(r[isamp])
    ops[op] += ex
    samp[op] += sa
    tot += ex
    tots += sa
print(f"total warp instructions {tot:.4g}, samples {tots}")
for op, c in ops.most_common(22):
    print(f"  {op:22s} {c:14d} {100.0*c/tot:5.1f}%  samples {100.0*samp[op]/max(tots,1):5.1f}%")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print("hottest lines by samples:")
idx = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:n]
for i in sorted(idx):
    r = body[i]
    print(f"  {i:5d} {r[isrc].strip()[:70]:70s} ex {int(r[iex]):12d} thr/w {int(r[ithr])/max(int(r[iex]),1):5.1f} samples {r[isamp]}")
