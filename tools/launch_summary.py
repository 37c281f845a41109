#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: tools/launch_summary.py <launches.csv> [out.csv]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "")
    v, u = float(r[mv].replace(",", "")), r[mu]
    tot[name] += v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    cnt[name] += 1
s = sum(tot.values())
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
out.write("kernel,launches,total_ms,share\n")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    out.write(f'"{k}",{cnt[k]},{v:.4f},{v / s:.4f}\n')
