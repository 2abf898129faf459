#!/usr/bin/env python3
"""Static opcode histogram of the built library's kernels (cuobjdump -sass), for profiles/:
    python tools/sass_hist.py [regex of demangled kernel names] > profiles/rNN_sass_opcodes.txt
Proves which hardware paths a kernel uses: UBLKCP (TMA bulk copy), SYNCS (mbarrier), IMAD.WIDE (64-bit multiply-add),
ATOMS (shared-memory atomics), RED/ATOMG (global atomics), MATCH/REDUX/VOTE (warp collectives), UTC*MMA (tcgen05)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "obake_b200", "lib", "libobk_b200.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_tileconv<3, false, true, 704|k_mul<1, 2, false>|k_planeconv<3, false, true, 8|k_rowconv<3, false, true|k_push|k_tile_units|k_bench_int_mac")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
cur, hist = None, None
res = []
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        if cur:
            res.append((cur, hist))
        cur, hist = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        hist[m.group(1)] += 1
if cur:
    res.append((cur, hist))
dem = subprocess.run(["c++filt"] + [r[0] for r in res], capture_output=True, text=True).stdout.splitlines()
print(f"# static SASS opcode counts, {os.path.relpath(LIB, ROOT)} (sm_100a), kernels matching /{pat.pattern}/")
for (mangled, hist), name in zip(res, dem):
    if not pat.search(name):
        continue
    tot = sum(hist.values())
    print(f"\n== {name[:150]}\n   {tot} instructions")
    key = [k for k in hist if re.match(r"UBLKCP|SYNCS|IMAD\.WIDE|ATOMS|ATOMG|RED|MATCH|REDUX|VOTE|UTC|LDS|STS|LDG|STG|SHFL|DFMA|DMUL|DADD|BAR|CREDUX|FENCE|MEMBAR", k)]
    print("   hardware paths: " + ", ".join(f"{k} x{hist[k]}" for k in sorted(key)))
    print("   top opcodes:    " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
