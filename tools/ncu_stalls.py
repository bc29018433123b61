#!/usr/bin/env python3
"""Per-instruction stall attribution from an .ncu-rep (source page): where do the samples of one stall reason sit?
usage: ncu_stalls.py rep kernel_regex [stall_column]"""
import csv, io, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
col = sys.argv[3] if len(sys.argv) > 3 else "stall_no_inst"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except ValueError: return 0.0
tot = sum(f(r, "# Samples") for r in data); totc = sum(f(r, col) for r in data); toti = sum(f(r, "Instructions Executed") for r in data)
print("instructions %d static, %.3e executed, samples %d, %s %d (%.1f%%)" % (len(data), toti, tot, col, totc, 100 * totc / max(tot, 1)))
# attribute to the instruction BEFORE (a stalled warp is sampled at the instruction it cannot issue; for fetch stalls the
# cause is usually the control transfer just before)
byop = Counter(); prevop = Counter()
for i, r in enumerate(data):
    t = r[ix["Source"]].split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    byop[op] += f(r, col)
    if i:
        t2 = data[i - 1][ix["Source"]].split()
        op2 = (t2[1] if t2 and t2[0].startswith("@") and len(t2) > 1 else (t2[0] if t2 else "?")).split(".")[0]
        prevop[op2] += f(r, col)
print("by stalled opcode:", [(k, int(v)) for k, v in byop.most_common(8)])
print("by preceding opcode:", [(k, int(v)) for k, v in prevop.most_common(8)])
# top addresses
top = sorted(range(len(data)), key=lambda i: -f(data[i], col))[:25]
for i in sorted(top):
    r = data[i]
    print("%5d %-60s %s=%d exec=%d  prev: %s" % (i, r[ix["Source"]].strip()[:60], col, f(r, col), f(r, "Instructions Executed"), data[i - 1][ix["Source"]].strip()[:40] if i else ""))
