import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bls_b200 import capi, hostgen as hg, layout as L
from oracle import pyoracle as orc
import __graft_entry__ as g
emu = ctypes.CDLL(g.build_emu())
U64 = np.uint64
_p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
def emu_op(op, arg, a):
    out = np.empty_like(a)
    emu.emu_quad_fp12_op(op, ctypes.c_uint64(arg), _p(a), _p(a), _p(out), None, ctypes.c_size_t(a.shape[0]))
    return out
ctx = capi.Ctx(0)
n = 8
P = hg.g1_progression(3, 1, n); Q = hg.g2_progression(4, 1, n)
f = orc.pairing_batch(P, Q).view(U64).reshape(n, 2, 3, 2, 6).copy()
for op, arg in [(20, 6), (20, 7), (20, 8), (20, 9), (20, 10), (20, 4)]:
    dev = ctx.test_op(4, op, f, None, arg)[0].reshape(n, 2, 3, 2, 6)
    ref = emu_op(op, arg, f)
    eq = (dev == ref).reshape(n, 6, -1).all(axis=2)
    print("op", op, "arg", hex(arg), "units ok", eq.all(axis=1).tolist(), "coeff ok (unit 0)", eq[0].tolist(), flush=True)
# which value does the device show in slot 4 of stage 6 (d[1] on half 1)?
dev6 = ctx.test_op(4, 20, f, None, 6)[0].reshape(n, 6, 2, 6)
emu6 = emu_op(20, 6, f).reshape(n, 6, 2, 6)
emu8 = emu_op(20, 8, f).reshape(n, 6, 2, 6)
u = 0
cands = {"emu d0": emu6[u, 0], "emu d1": emu6[u, 4], "emu d2": emu8[u, 4], "emu C0.g2": emu6[u, 3], "emu C0.g3": emu6[u, 2], "emu C0.g4": emu6[u, 1], "emu C0.g5": emu6[u, 5]}
for k, v in cands.items():
    print(k, "== dev slot4:", bool((dev6[u, 4] == v).all()), " j0:", bool((dev6[u, 4, 0] == v[0]).all()), " j1:", bool((dev6[u, 4, 1] == v[1]).all()))
print("dev slot 4 j0 == emu d1 j0", bool((dev6[u, 4, 0] == emu6[u, 4, 0]).all()), "j1", bool((dev6[u, 4, 1] == emu6[u, 4, 1]).all()))
print("dev slot4", [hex(L.limbs_to_int(dev6[u, 4, j])) for j in range(2)])
print("emu slot4", [hex(L.limbs_to_int(emu6[u, 4, j])) for j in range(2)])
