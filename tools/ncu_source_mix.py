#!/usr/bin/env python3
"""Instruction mix / hot regions from `ncu -i rep --page source --csv` output (csv path as argv[1])."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Instructions Executed" in r)
isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ist = hdr.index("# Samples"); iavg = hdr.index("Avg. Threads Executed")
def f(x):
    try: return float(x)
    except ValueError: return None
data = [r for r in rows if len(r) > iavg and f(r[iex]) is not None]
tot = sum(f(r[iex]) for r in data)
print("total warp instr %.3e, sass lines %d" % (tot, len(data)))
c = Counter(); s = Counter()
for r in data:
    t = r[isrc].split()
    if not t: continue
    op = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
    c[op] += f(r[iex]); s[op] += f(r[ist]) or 0
ss = sum(s.values()) or 1
for op, v in c.most_common(22):
    print("%-10s %6.2f%% instr  %6.2f%% samples" % (op, 100 * v / tot, 100 * s[op] / ss))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 150
for k in range(0, len(data), step):
    seg = data[k:k + step]
    print(k, "%.2f%% instr" % (100 * sum(f(r[iex]) for r in seg) / tot), "avgthr %.1f" % (sum(f(r[iavg]) or 0 for r in seg) / len(seg)),
          "samples %.2f%%" % (100 * sum(f(r[ist]) or 0 for r in seg) / ss), seg[0][isrc][:60])
