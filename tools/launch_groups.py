#!/usr/bin/env python3
"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list into calls (a call starts at the kernel given as argv[2])
and print per-kernel times of the last few calls: python tools/launch_groups.py launches.csv k_msm_hist [ncalls]"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
seq = [(r[kn].split('(')[0].replace('void b381::', '').replace('b381::', '').replace('void ', ''), float(r[mv].replace(',', '')) / 1e6) for r in rows[hi + 1:] if len(r) > mv]
calls = []; cur = []
for k, t in seq:
    if k.startswith(sys.argv[2]) and cur: calls.append(cur); cur = []
    cur.append((k, t))
calls.append(cur)
for c in calls[-int(sys.argv[3]) if len(sys.argv) > 3 else -4:]:
    d = OrderedDict()
    for k, t in c: d[k] = d.get(k, 0) + t
    print("%.2f ms: " % sum(d.values()) + ", ".join("%s %.2f" % (k[:30], v) for k, v in d.items()))
