"""The Fq2-granular VM programs of the pairing path (see trace.py / sched.py / csrc/vm.cuh).

Segments of the per-unit global area: 0 and 1 inputs, 2 output, 3 spill, 4 constants.  Indices are in Fq.
  ml    : seg0 = G1 affine (x, y), seg1 = G2 affine (x, y as Fq2) -> seg2 = Fq12 (6 Fq2)
  fe_a  : seg0 = Fq12 f -> seg2 = the Fq norm n whose inverse the Fq12 inversion needs (1 Fq)
  fe_c  : seg0 = Fq12 f, seg1 = n^-1 (1 Fq) -> seg2 = FinalExponentiation(f)
"""
from . import trace as T
from . import sched as S


def frobenius_tables(p):
    Qm = T.Q

    def f2mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Qm, (a[0] * b[1] + a[1] * b[0]) % Qm)

    def f2pow(a, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = f2mul(r, a)
            a = f2mul(a, a)
            e >>= 1
        return r
    tabs = {"fq6_c1": {}, "fq6_c2": {}, "fq12_c1": {}}
    for k in (1, 2, 3):
        for name, d, mult in (("fq6_c1", 3, 1), ("fq6_c2", 3, 2), ("fq12_c1", 6, 1)):
            c = f2pow((1, 1), mult * (Qm ** k - 1) // d)
            tabs[name][k] = p.const(c[0], c[1])
    return tabs


def build_ml(npairs=1):
    p = T.Program("ml%d" % npairs)
    pairs = []
    for i in range(npairs):
        px, py = p.load(0, 2 * i, 1), p.load(0, 2 * i + 1, 1)
        qx, qy = p.load(1, 4 * i), p.load(1, 4 * i + 2)
        pairs.append((px, py, qx, qy))
    f = T.miller_loop(p, pairs)
    for i, c in enumerate(f.coeffs()):
        p.store(c, 2, 2 * i)
    return p


def _load_f(p):
    return T.fq12_from([p.load(0, 2 * i) for i in range(6)])


def build_fe_a():
    p = T.Program("fe_a")
    n, _ = T.fq12_inv_norm(_load_f(p))
    p.store(n, 2, 0, 1)
    return p


def build_fe_c(spill=True):
    p = T.Program("fe_c")
    f = _load_f(p)
    ninv = p.load(1, 0, 1)
    tabs = frobenius_tables(p)
    _, inter = T.fq12_inv_norm(f)

    def sp(v):
        return T.fq12_from([p.spill(c) for c in v.coeffs()])
    out = T.final_exp(f, ninv, inter, tabs, sp if spill else None)
    for i, c in enumerate(out.coeffs()):
        p.store(c, 2, 2 * i)
    return p


def compile_program(p, L, window=0, do_fold=True):
    ops = S.fold(p.ops)
    steps = S.schedule(ops, L, window)
    slot, nslots = S.allocate(ops, steps)
    code = S.encode(ops, steps, slot, L)
    return {"name": p.name, "L": L, "code": code, "nslots": nslots, "consts": list(p.consts), "nsteps": len(steps),
            "spill_fq": p.spill_fq, "stats": S.stats(ops, steps, L)}
