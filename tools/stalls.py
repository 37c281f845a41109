#!/usr/bin/env python
"""Stall-reason breakdown (sampled) of one kernel in an ncu report.  usage: tools/stalls.py <rep> <kernel> [idx]"""
import csv, subprocess, sys, io
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if 'Instructions Executed' in r]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = his[which]; end = his[which + 1] - 1 if which + 1 < len(his) else len(rows)
hdr = rows[hi]
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = {}
for r in rows[hi + 1:end]:
    for i in stall:
        try: tot[hdr[i]] = tot.get(hdr[i], 0) + float(r[i])
        except Exception: pass
s = sum(tot.values())
print(sys.argv[2], 'samples', int(s), sorted([(round(v / s, 3), k) for k, v in tot.items() if v / s > 0.02], reverse=True))
