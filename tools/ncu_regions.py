#!/usr/bin/env python3
"""Contiguous hot regions of a kernel's SASS from an .ncu-rep: size, executed share, opcode mix, stall shares.
usage: ncu_regions.py rep kernel_regex"""
import csv, io, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except ValueError: return 0.0
ex = [f(r, "Instructions Executed") for r in data]
tot = sum(ex); tots = sum(f(r, "# Samples") for r in data)
def op(r):
    t = r[ix["Source"]].split()
    return (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
# split at RET / unconditional BRA / EXIT boundaries = function-ish regions
regs = []; start = 0
for i, r in enumerate(data):
    o = op(r)
    if o in ("RET", "EXIT") or (o == "BRA" and not r[ix["Source"]].strip().startswith("@")):
        regs.append((start, i + 1)); start = i + 1
if start < len(data): regs.append((start, len(data)))
# merge tiny regions into neighbours
out = []
for a, b in regs:
    e = sum(ex[a:b])
    if e / tot < 0.002: continue
    c = Counter(op(r) for r in data[a:b])
    smp = sum(f(r, "# Samples") for r in data[a:b])
    ni = sum(f(r, "stall_no_inst") for r in data[a:b]); ls = sum(f(r, "stall_long_sb") for r in data[a:b]); w = sum(f(r, "stall_wait") for r in data[a:b])
    out.append((e, a, b, c, smp, ni, ls, w))
out.sort(reverse=True)
print("%-6s %-6s %7s %7s %7s %6s %6s %6s  mix" % ("start", "instrs", "KB", "exec%", "smp%", "noinst", "longsb", "wait"))
for e, a, b, c, smp, ni, ls, w in out[:40]:
    print("%-6d %-6d %7.1f %7.2f %7.2f %6.1f %6.1f %6.1f  %s" % (a, b - a, (b - a) * 16 / 1024, 100 * e / tot, 100 * smp / tots, 100 * ni / max(smp, 1), 100 * ls / max(smp, 1),
          100 * w / max(smp, 1), " ".join("%s:%d" % kv for kv in c.most_common(5))))
# hot set: how much static code holds 50 / 80 / 90 / 95 / 99 % of the executed instructions (the instruction-cache question)
order = sorted(range(len(ex)), key=lambda i: -ex[i]); cum = 0.0; marks = [0.5, 0.8, 0.9, 0.95, 0.99]; mi = 0
print("static instructions %d (%.1f KB), executed %.3e" % (len(ex), len(ex) * 16 / 1024, tot))
for k, i in enumerate(order):
    cum += ex[i]
    while mi < len(marks) and cum >= marks[mi] * tot:
        print("hot set: %2.0f %% of the executed instructions in %.1f KB" % (100 * marks[mi], (k + 1) * 16 / 1024)); mi += 1
