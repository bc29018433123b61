#!/usr/bin/env python3
"""GPU spot check of one kernel path against the oracle (small ragged batches, infinity pairs, zero final-exp input):
python tools/quad_check.py [path]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bls_b200 import capi, hostgen as hg, layout as L
from oracle import pyoracle as orc
path = sys.argv[1] if len(sys.argv) > 1 else "quad"
ctx = capi.Ctx(0, path=path)
for n in (1, 7, 33, 300):
    P = hg.g1_progression(0xB2000002 + n, 0x1234567, n); Q = hg.g2_progression(0x5EED + n, 0x7654321, n)
    if n > 5: P["inf"][3] = 1; Q["inf"][5] = 1
    got = ctx.pairing_batch(P, Q)
    P2 = P.copy(); Q2 = Q.copy(); P2["inf"][:] = 0; Q2["inf"][:] = 0
    exp = orc.pairing_batch(P2, Q2, threads=8)
    one = np.zeros((2, 3, 2, 6), np.uint64); one[0, 0, 0] = L.fp_from_int(1)
    bad = 0
    for i in range(n):
        w = one if (P["inf"][i] or Q["inf"][i]) else exp[i].view(np.uint64).reshape(2, 3, 2, 6)
        if not (got[i].view(np.uint64).reshape(2, 3, 2, 6) == w).all(): bad += 1
    print(path, "n", n, "mismatches", bad, flush=True)
    assert bad == 0
ml = ctx.miller_loop_batch(P, Q)
for i in (0, 1, 299):
    assert (ml[i] == orc.miller_loop(P[i:i + 1], Q[i:i + 1])).all(), i
xs = orc.XorShift(3)
f = xs.rand_fq(12 * 9).reshape(9, 2, 3, 2, 6); f[4] = 0
fe, ok = ctx.final_exp_batch(f)
assert ok.tolist() == [1, 1, 1, 1, 0, 1, 1, 1, 1]
for i in range(9):
    good, e = orc.final_exp(f[i])
    if good: assert (fe[i] == e).all(), i
nP = P[:1].copy(); nP["y"][0] = orc.fq("neg", nP["y"][0])[0]
okk = ctx.pairing_product_is_one(np.concatenate([P[:1], nP, P[:1], P[1:2]]), np.concatenate([Q[:1], Q[:1], Q[:1], Q[:1]]), [0, 2, 4])
assert okk.tolist() == [1, 0], okk
print(path, "checks ok")
