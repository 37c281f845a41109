#!/usr/bin/env python
"""Summarise an ncu report's SASS source page for one kernel into hot regions (runs of instructions
with similar execution counts): share of executed instructions, share of stall samples, opcode mix.
usage: tools/sass_hot.py <report.ncu-rep> <kernel-name> [n_regions]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r]
hdr = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
ii, ss, si = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
def f(x):
    try: return float(x)
    except Exception: return None
data = [(r[si].strip(), f(r[ii]), f(r[ss]) or 0) for r in rows[hi[0] + 1:end] if len(r) > ii and f(r[ii]) is not None]
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print(len(data), 'sass instr; warp-instr executed', tot, 'samples', tots)
seg = []; cur = None
for i, (s, e, sm) in enumerate(data):
    if cur is None or not (0.5 * cur['e'] <= e <= 2 * cur['e']):
        cur = {'start': i, 'e': e if e > 0 else 1, 'n': 0, 'sum': 0, 'smp': 0, 'ops': {}}; seg.append(cur)
    cur['n'] += 1; cur['sum'] += e; cur['smp'] += sm
    t = s.split(); op = t[0] if not t[0].startswith('@') else t[1]
    op = op.split('.')[0]; cur['ops'][op] = cur['ops'].get(op, 0) + 1
for sg in sorted(seg, key=lambda x: -x['sum'])[:top]:
    ops = sorted(sg['ops'].items(), key=lambda x: -x[1])[:8]
    print('start %5d n %5d exec/instr %.3g share %.3f samples %.3f' % (sg['start'], sg['n'], sg['sum'] / sg['n'], sg['sum'] / tot, sg['smp'] / tots), ops)
