"""The VM programs of the pairing path and their (deterministic) code generation.

Segments of the per-unit global area (kernel arguments, see vm.cuh): 0 and 1 inputs, 2 output, 3 spill.
  ml    : seg0 = G1 affine (x, y), seg1 = G2 affine (x.c0, x.c1, y.c0, y.c1) -> seg2 = Fq12 (12 Fq)
  fe_a  : seg0 = Fq12 f -> seg2 = the Fq norm n whose inverse the Fq12 inversion needs (1 Fq)
  fe_c  : seg0 = Fq12 f, seg1 = n^-1 (1 Fq) -> seg2 = FinalExponentiation(f) (12 Fq)
"""
from . import trace as T
from . import sched as S


def frobenius_tables(p):
    """(1+u)^((Q^k - 1)/d) as Fq2 of constants: fq6.go:144-208, fq12.go:122-168 (regenerated)"""
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
    xi = (1, 1)
    tabs = {"fq6_c1": {}, "fq6_c2": {}, "fq12_c1": {}}
    for k in (1, 2, 3):
        for name, d, mult in (("fq6_c1", 3, 1), ("fq6_c2", 3, 2), ("fq12_c1", 6, 1)):
            c = f2pow(xi, mult * (Qm ** k - 1) // d)
            tabs[name][k] = T.Fq2(p.const(c[0]), p.const(c[1]))
    return tabs


def build_ml(npairs=1):
    p = T.Program("ml%d" % npairs)
    pairs = []
    for i in range(npairs):
        px, py = p.load(0, 2 * i), p.load(0, 2 * i + 1)
        q = [p.load(1, 4 * i + j) for j in range(4)]
        pairs.append((px, py, T.Fq2(q[0], q[1]), T.Fq2(q[2], q[3])))
    f = T.miller_loop(p, pairs)
    for i, c in enumerate(f.coeffs()):
        p.output(c, 2, i)
    return p


def _load_f(p):
    return T.fq12_from([p.load(0, i) for i in range(12)])


def build_fe_a():
    p = T.Program("fe_a")
    n, _ = T.fq12_inv_norm(_load_f(p))
    p.output(n, 2, 0)
    return p


def build_fe_c(spill=True):
    p = T.Program("fe_c")
    f = _load_f(p)
    ninv = p.load(1, 0)
    tabs = frobenius_tables(p)
    _, inter = T.fq12_inv_norm(f)
    nspill = [0]

    def sp(v):
        cs = []
        for c in v.coeffs():
            cs.append(p.store(c, 3, nspill[0])); nspill[0] += 1
        return T.fq12_from(cs)
    out = T.final_exp(f, ninv, inter, tabs, sp if spill else None)
    for i, c in enumerate(out.coeffs()):
        p.output(c, 2, i)
    p.spill_fq = nspill[0]
    return p


def compile_program(p, L, window=0):
    steps = S.schedule(p, L, window)
    slot, nslots = S.allocate(p, steps)
    code, padded = S.encode(p, steps, slot, L)
    return {"name": p.name, "L": L, "code": code, "nslots": nslots, "consts": list(p.consts), "nsteps": len(steps),
            "spill_fq": getattr(p, "spill_fq", 0), "stats": S.stats(p, steps, L, padded)}
