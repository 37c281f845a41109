#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS in a built library (no GPU needed).
usage: tools/sass_ops.py <lib.so> <kernel-name-substring> [top]"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ops = None, collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        ops[cur][m.group(1)] += 1
for fn, c in ops.items():
    if pat in fn:
        print(fn, "total", sum(c.values()))
        print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(top)))
