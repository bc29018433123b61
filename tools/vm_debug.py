#!/usr/bin/env python3
"""Micro-programs through b381_vm_exec_dev vs the emulator (debugging aid for csrc/vm.cuh)."""
import ctypes, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bls_b200 import capi, layout as L
from bls_b200.vm import sched as S, trace as T
ctx = capi.Ctx(0)
rng = random.Random(1)
nin = 3
def run(name, fmas):
    p = T.Program("dbg")
    vals = [p.load(0, 2 * i).single()[0] for i in range(nin)]
    ops = list(p.ops)
    outs = []
    for f in fmas:
        d = p._new()
        op = {"kind": "FMA", "mode": f.get("mode", "lin"), "dst": d, "a": [(vals[i], s, x, c) for i, s, x, c in f.get("a", [])],
              "b": [(vals[i], s, x, c) for i, s, x, c in f.get("b", [])], "g1": [(vals[i], s, x, c) for i, s, x, c in f.get("g1", [])],
              "m3": f.get("m3", 0), "add": [(vals[i], s, x, c) for i, s, x, c in f.get("add", [])], "pxi": f.get("pxi", 0)}
        ops.append(op); vals.append(d); outs.append(d)
    if not fmas:
        outs = [vals[1]]
    for i, v in enumerate(outs):
        ops.append({"kind": "IO", "op": "store", "src": v, "seg": 2, "idx": 2 * i, "width": 2, "dst_handle": -1000 - i})
    steps = S.schedule(ops, 4, 0); slot, nslots = S.allocate(ops, steps); code = S.encode(ops, steps, slot, 4)
    nunits, nout = 9, len(outs)
    ints = [[rng.randrange(T.Q) for _ in range(2 * nin)] for _ in range(nunits)]
    ins = np.array([[L.int_to_limbs(v) for v in row] for row in ints], np.uint64)
    d_in = ctx.to_device(ins); d_out = ctx.dev_empty(nunits * 2 * nout * 48); d_x = ctx.dev_empty(64)
    segs = (ctypes.c_void_p * 4)(d_in.ptr.value, d_x.ptr.value, d_out.ptr.value, d_x.ptr.value)
    strides = (ctypes.c_size_t * 4)(2 * nin * 48, 0, 2 * nout * 48, 0)
    ctx.call("b381_vm_exec_dev", ctypes.c_char_p(code), 4, len(steps), nslots, None, 0, segs, strides, ctypes.c_size_t(nunits))
    got = ctx.from_device(d_out, np.uint64, nunits * 2 * nout * 6).reshape(nunits, 2 * nout, 6)
    emu = S.Emulator(code, 4, nslots, [])
    bad = 0
    for u in range(nunits):
        exp = emu.run({0: list(ints[u]), 1: [], 2: [], 3: []})[2]
        gotu = [L.limbs_to_int(x) for x in got[u]]
        bad += sum(1 for a, b in zip(gotu, exp) if a != b)
    print("%-28s %s (%d steps, %d mismatching coefficients of %d)" % (name, "ok" if not bad else "FAIL", len(steps), bad, nunits * 2 * nout))
    if bad and os.environ.get("VERBOSE"):
        exp = emu.run({0: list(ints[0]), 1: [], 2: [], 3: []})[2]
        print("   in  ", [hex(v)[:20] for v in ints[0]])
        print("   got ", [hex(L.limbs_to_int(x))[:20] for x in got[0]])
        print("   exp ", [hex(v)[:20] for v in exp])
        print("   got-exp mod Q", [hex((L.limbs_to_int(x) - e) % T.Q)[:20] for x, e in zip(got[0], exp)])
P = lambda i: (i, 1, 0, 0)
run("io only", [])
run("copy", [{"add": [P(0)]}])
run("add2", [{"add": [P(0), P(1)]}])
run("sub", [{"add": [P(0), (1, -1, 0, 0)]}])
run("neg only", [{"add": [(1, -1, 0, 0)]}])
run("xi", [{"add": [(0, 1, 1, 0)]}])
run("conj", [{"add": [(0, 1, 0, 1)]}])
run("neg xi conj", [{"add": [(0, -1, 1, 1), (2, 1, 0, 0)]}])
run("8 terms", [{"add": [P(0), P(1), P(2), (0, -1, 0, 0), (1, -1, 1, 0), P(2), P(2), (1, 1, 0, 1)]}])
run("m3 lin", [{"g1": [P(0), (1, -1, 0, 0)], "m3": 1, "add": [P(2), P(2)]}])
run("mul plain", [{"mode": "mul", "a": [P(0)], "b": [P(1)]}])
run("sqr plain", [{"mode": "sqr", "a": [P(0)]}])
run("mul operands", [{"mode": "mul", "a": [P(0), (1, -1, 0, 0)], "b": [P(1), P(2), (0, 1, 1, 0)]}])
run("mul pxi add", [{"mode": "mul", "a": [P(0)], "b": [P(1)], "pxi": 1, "add": [(2, -1, 0, 0)]}])
run("sqr m3", [{"mode": "sqr", "a": [P(0), P(1)], "g1": [(2, -1, 0, 0)], "m3": 1, "add": [P(1), P(1)]}])
run("two ops", [{"mode": "mul", "a": [P(0)], "b": [P(1)]}, {"mode": "sqr", "a": [P(2)]}])
