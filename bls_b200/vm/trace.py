"""Symbolic tracer for the warp-cooperative Fq virtual machine (bls_b200/csrc/vm.cuh).

The pairing is expressed once, in Python, over symbolic Fq values; tracing it yields a DAG of two
kinds of operations -- MUL (one 384-bit Montgomery multiplication, operands optionally formed as
x +/- y on the fly) and LIN (a signed sum of up to 8 terms with power-of-two multipliers, reduced to
the canonical residue) -- which sched.py packs into warp-synchronous steps.

Additions are *lazy*: an `Fq` is a linear combination {value: small integer}; nothing is emitted
until a combination is needed as a multiplication operand or as an output, so Karatsuba
recombinations collapse into a few wide LIN ops instead of long chains of two-term additions.

The formulas are the ones of bls_b200/csrc/{tower,pairing}.cuh (themselves citing the reference:
fq2.go, fq6.go, fq12.go, g2.go:655-772, pairing.go:16-129); all values are canonical residues, so
the results are bit-identical to the reference's whatever the schedule.
"""
from __future__ import annotations

MAX_WEIGHT = 8          # sum of |multipliers| in one LIN op: the accumulator stays below 8Q < 2^384
LAZY_TERMS = 4          # combinations with more terms than this are materialised when created


CONST_SEG = 4            # global segment of the constant table (stride 0: shared by all units)
RELOAD_AFTER = 250       # ops after which a cached copy of a global/constant value is loaded afresh


class Program:
    """ops: MUL {dst, a: [(v, sign)], b: [...]}, LIN {dst, terms: [(v, sign)]}, IO {dst|src, op, seg, idx}.
    MUL / LIN operands are always shared-memory slot values; constants, inputs and spilled values reach
    a slot through an IO load (and leave through an IO store)."""

    def __init__(self, name):
        self.name = name
        self.ops = []
        self.nvals = 0
        self.consts = []         # canonical integers (NOT Montgomery) of the constant table
        self._const_ids = {}
        self._lin_cache = {}
        self._glob = {}          # global handle -> (seg, idx)
        self._glob_copy = {}     # global handle -> (slot value, op count when loaded)
        self.outputs = []

    # -- values ---------------------------------------------------------------------------------
    def _new(self):
        v = self.nvals
        self.nvals += 1
        return v

    def _glob_handle(self, seg, idx):
        h = -(len(self._glob) + 1)          # negative ids: values that live in global memory
        self._glob[h] = (seg, idx)
        return h

    def const(self, integer):
        integer %= Q
        if integer not in self._const_ids:
            self._const_ids[integer] = self._glob_handle(CONST_SEG, len(self.consts))
            self.consts.append(integer)
        return Fq(self, {self._const_ids[integer]: 1})

    def glob(self, seg, idx):
        return Fq(self, {self._glob_handle(seg, idx): 1})

    def load(self, seg, idx):
        """input -> slot value"""
        return Fq(self, {self._slot_of(self._glob_handle(seg, idx)): 1})

    def _slot_of(self, v):
        """slot value holding v (v itself, or a recent IO load of a global value)"""
        if v >= 0:
            return v
        c = self._glob_copy.get(v)
        if c is None or len(self.ops) - c[1] > RELOAD_AFTER:
            d = self._new()
            seg, idx = self._glob[v]
            self.ops.append({"kind": "IO", "op": "load", "dst": d, "seg": seg, "idx": idx, "src_handle": v})
            c = (d, len(self.ops))
            self._glob_copy[v] = c
        return c[0]

    def store(self, x, seg, idx):
        """x -> global (seg, idx); returns the global-resident value (re-loaded on later use)"""
        m = self.materialise(x)
        h = self._glob_handle(seg, idx)
        self.ops.append({"kind": "IO", "op": "store", "src": m.single(), "seg": seg, "idx": idx, "dst_handle": h})
        return Fq(self, {h: 1})

    # -- emission -------------------------------------------------------------------------------
    def _emit_lin(self, dst, terms):
        """terms: dict value -> integer coefficient (expanded into repeated +/-1 terms);
        splits into chunks of weight <= MAX_WEIGHT"""
        flat = []            # (value, sign)
        for v, c in terms.items():
            flat += [(v, 1 if c > 0 else -1)] * abs(c)
        if not flat:                       # the zero combination
            flat = [(self.const(0).single(), 1)]
        while len(flat) > MAX_WEIGHT:
            chunk, rest = flat[:MAX_WEIGHT], flat[MAX_WEIGHT:]
            tmp = self._new()
            self.ops.append({"kind": "LIN", "dst": tmp, "terms": [(self._slot_of(v), sg) for v, sg in chunk]})
            flat = rest + [(tmp, 1)]
        self.ops.append({"kind": "LIN", "dst": dst, "terms": [(self._slot_of(v), sg) for v, sg in flat]})

    def materialise(self, x):
        """Fq -> Fq that is a single slot value with coefficient +1"""
        if x.is_single():
            return x
        if len(x.terms) == 1 and next(iter(x.terms.values())) == 1:      # a bare global / constant
            return Fq(self, {self._slot_of(x.single()): 1})
        key = tuple(sorted(x.terms.items()))
        if key in self._lin_cache:
            return self._lin_cache[key]
        dst = self._new()
        self._emit_lin(dst, x.terms)
        r = Fq(self, {dst: 1})
        self._lin_cache[key] = r
        return r

    def _operand(self, x):
        """-> ([(slot value, sign)] with 1-2 entries and a leading +, overall sign)"""
        t = [(v, c) for v, c in x.terms.items() if c]
        if len(t) == 0:
            return None, 0
        if len(t) > 2 or any(abs(c) != 1 for _, c in t):
            x = self.materialise(x)
            t = list(x.terms.items())
        t = [(self._slot_of(v), c) for v, c in t]
        if len(t) == 1:
            return [(t[0][0], 1)], (1 if t[0][1] > 0 else -1)
        (v0, c0), (v1, c1) = t
        if c0 > 0 and c1 > 0:
            return [(v0, 1), (v1, 1)], 1
        if c0 < 0 and c1 < 0:
            return [(v0, 1), (v1, 1)], -1
        if c0 > 0:
            return [(v0, 1), (v1, -1)], 1
        return [(v1, 1), (v0, -1)], 1

    def _is_const(self, x, integer):
        if len(x.terms) != 1:
            return False
        (v, c), = x.terms.items()
        return c == 1 and v < 0 and self._glob[v][0] == CONST_SEG and self.consts[self._glob[v][1]] == integer

    def mul(self, a, b):
        return self.dot([(a, b)])

    def dot(self, pairs):
        """sum of products a_i * b_i with ONE Montgomery reduction per two products (lazy reduction of
        the Fq2 rows: fq2.go:116-130 computes the same values with three reduced multiplications)"""
        extra = Fq(self, {})
        prods = []
        for a, b in pairs:
            if self._is_const(a, 0) or self._is_const(b, 0) or not a.terms or not b.terms:
                continue
            if self._is_const(a, 1):
                extra = extra + b
                continue
            if self._is_const(b, 1):
                extra = extra + a
                continue
            oa, sa = self._operand(a)
            ob, sb = self._operand(b)
            prods.append((oa, ob, sa * sb))
        res = extra
        for i in range(0, len(prods), 2):
            grp = prods[i:i + 2]
            flip = grp[0][2] < 0
            operands = []
            for oa, ob, sg in grp:
                if flip:
                    sg = -sg
                if sg < 0:
                    ob = [(v, -t) for v, t in ob]
                operands += [oa, ob]
            dst = self._new()
            self.ops.append({"kind": "MUL", "k": len(grp), "dst": dst, "operands": operands})
            res = res + Fq(self, {dst: -1 if flip else 1})
        return res

    def output(self, x, seg, idx):
        self.outputs.append(self.store(x, seg, idx))


Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


class Fq:
    __slots__ = ("p", "terms")

    def __init__(self, p, terms):
        self.p = p
        self.terms = {v: c for v, c in terms.items() if c}
        if len(self.terms) > LAZY_TERMS or sum(abs(c) for c in self.terms.values()) > 2 * MAX_WEIGHT:
            m = p.materialise(self)
            self.terms = dict(m.terms)

    def is_single(self):
        return len(self.terms) == 1 and next(iter(self.terms.values())) == 1 and next(iter(self.terms)) >= 0

    def single(self):
        assert len(self.terms) == 1
        return next(iter(self.terms))

    def _comb(self, o, k):
        t = dict(self.terms)
        for v, c in o.terms.items():
            t[v] = t.get(v, 0) + k * c
        return Fq(self.p, t)

    def __add__(self, o): return self._comb(o, 1)
    def __sub__(self, o): return self._comb(o, -1)
    def __neg__(self): return Fq(self.p, {v: -c for v, c in self.terms.items()})
    def scale(self, k): return Fq(self.p, {v: k * c for v, c in self.terms.items()})
    def dbl(self): return self.scale(2)
    def __mul__(self, o): return self.p.mul(self, o)
    def sqr(self): return self.p.mul(self, self)


# ---- tower: tuples of Fq -------------------------------------------------------------------------
class Fq2:
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def __add__(self, o): return Fq2(self.c0 + o.c0, self.c1 + o.c1)
    def __sub__(self, o): return Fq2(self.c0 - o.c0, self.c1 - o.c1)
    def __neg__(self): return Fq2(-self.c0, -self.c1)
    def dbl(self): return Fq2(self.c0.dbl(), self.c1.dbl())
    def scale(self, k): return Fq2(self.c0.scale(k), self.c1.scale(k))
    def conj(self): return Fq2(self.c0, -self.c1)
    def mul_nr(self): return Fq2(self.c0 - self.c1, self.c0 + self.c1)          # * (1 + u), fq2.go:41-45

    def __mul__(self, o):                       # fq2.go:116-130: same value, rows as two-product dot products
        p = self.c0.p
        return Fq2(p.dot([(self.c0, o.c0), (self.c1, -o.c1)]), p.dot([(self.c0, o.c1), (self.c1, o.c0)]))

    def sqr(self):                                                              # fq2.go:75-89
        s = (self.c0 + self.c1) * (self.c0 - self.c1)
        return Fq2(s, self.c0 * (self.c1 + self.c1))

    def mul_fq(self, k): return Fq2(self.c0 * k, self.c1 * k)
    def mat(self):
        p = self.c0.p
        return Fq2(p.materialise(self.c0), p.materialise(self.c1))
    def coeffs(self): return [self.c0, self.c1]


class Fq6:
    __slots__ = ("c0", "c1", "c2")

    def __init__(self, c0, c1, c2):
        self.c0, self.c1, self.c2 = c0, c1, c2

    def __add__(self, o): return Fq6(self.c0 + o.c0, self.c1 + o.c1, self.c2 + o.c2)
    def __sub__(self, o): return Fq6(self.c0 - o.c0, self.c1 - o.c1, self.c2 - o.c2)
    def __neg__(self): return Fq6(-self.c0, -self.c1, -self.c2)
    def dbl(self): return Fq6(self.c0.dbl(), self.c1.dbl(), self.c2.dbl())
    def mul_nr(self): return Fq6(self.c2.mul_nr(), self.c0, self.c1)            # * v, fq6.go:34-37

    def __mul__(self, o):                                                       # fq6.go:255-292
        a, b = self, o
        v0, v1, v2 = a.c0 * b.c0, a.c1 * b.c1, a.c2 * b.c2
        x = ((a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2).mul_nr() + v0
        y = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1 + v2.mul_nr()
        z = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2 + v1
        return Fq6(x, y, z)

    def mul_by_01(self, b0, b1):                                                # fq6.go:60-90
        a = self
        v0, v1 = a.c0 * b0, a.c1 * b1
        x = ((a.c1 + a.c2) * b1 - v1).mul_nr() + v0
        y = (a.c0 + a.c1) * (b0 + b1) - v0 - v1
        z = (a.c0 + a.c2) * b0 - v0 + v1
        return Fq6(x, y, z)

    def mul_by_1(self, b1):                                                     # fq6.go:40-57
        return Fq6((self.c2 * b1).mul_nr(), self.c0 * b1, self.c1 * b1)

    def frobenius(self, power, tabs):                                           # fq6.go:211-218
        c0, c1, c2 = self.c0, self.c1, self.c2
        if power & 1:
            c0, c1, c2 = c0.conj(), c1.conj(), c2.conj()
        return Fq6(c0, c1 * tabs["fq6_c1"][power], c2 * tabs["fq6_c2"][power])

    def mat(self): return Fq6(self.c0.mat(), self.c1.mat(), self.c2.mat())
    def coeffs(self): return self.c0.coeffs() + self.c1.coeffs() + self.c2.coeffs()


class Fq12:
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def conj(self): return Fq12(self.c0, -self.c1)                              # fq12.go:27-29

    def __mul__(self, o):                                                       # fq12.go:198-213
        aa = self.c0 * o.c0
        bb = self.c1 * o.c1
        s = (self.c0 + self.c1) * (o.c0 + o.c1)
        return Fq12(bb.mul_nr() + aa, s - aa - bb)

    def sqr(self):                                                              # fq12.go:180-195
        a, b = self.c0, self.c1
        ab = a * b
        s = (a + b) * (b.mul_nr() + a)
        return Fq12(s - ab - ab.mul_nr(), ab.dbl())

    def mul_by_014(self, d0, d1, d4):                                           # fq12.go:32-47
        aa = self.c0.mul_by_01(d0, d1)
        bb = self.c1.mul_by_1(d4)
        s = (self.c1 + self.c0).mul_by_01(d0, d1 + d4)
        return Fq12(bb.mul_nr() + aa, s - aa - bb)

    def frobenius(self, power, tabs):                                           # fq12.go:171-177
        c0 = self.c0.frobenius(power, tabs)
        c1 = self.c1.frobenius(power, tabs)
        k = tabs["fq12_c1"][power]
        return Fq12(c0, Fq6(c1.c0 * k, c1.c1 * k, c1.c2 * k))

    def cyclotomic_sqr(self):                                                   # Granger-Scott, tower.cuh
        def fp4(a, b):
            t0, t1 = a.sqr(), b.sqr()
            s = (a + b).sqr()
            return t0 + t1.mul_nr(), s - t0 - t1
        z0, z4, z3, z2, z1, z5 = self.c0.c0, self.c0.c1, self.c0.c2, self.c1.c0, self.c1.c1, self.c1.c2
        t0, t1 = fp4(z0, z1)
        n00 = t0.scale(3) - z0.dbl()
        n11 = t1.scale(3) + z1.dbl()
        t0, t1 = fp4(z2, z3)
        t2, t3 = fp4(z4, z5)
        n01 = t0.scale(3) - z4.dbl()
        n12 = t1.scale(3) + z5.dbl()
        t3 = t3.mul_nr()
        n10 = t3.scale(3) + z2.dbl()
        n02 = t2.scale(3) - z3.dbl()
        return Fq12(Fq6(n00, n01, n02), Fq6(n10, n11, n12))

    def mat(self): return Fq12(self.c0.mat(), self.c1.mat())
    def coeffs(self): return self.c0.coeffs() + self.c1.coeffs()


def fq12_from(coeffs):
    c = list(coeffs)
    f2 = [Fq2(c[2 * i], c[2 * i + 1]) for i in range(6)]
    return Fq12(Fq6(f2[0], f2[1], f2[2]), Fq6(f2[3], f2[4], f2[5]))


def fq12_one(p):
    z = p.const(0)
    return fq12_from([p.const(1)] + [z] * 11)


# ---- inversion down to one Fq: norm chain (fq12.go:216-237, fq6.go:295-336, fq2.go:133-147) --------
def fq12_inv_norm(f):
    """the Fq element whose inverse the Fq12 inversion needs, plus the intermediates"""
    t0 = f.c0 * f.c0 - (f.c1 * f.c1).mul_nr()          # Fq6
    a = t0
    k0 = a.c0.sqr() - (a.c1 * a.c2).mul_nr()
    k1 = a.c2.sqr().mul_nr() - a.c0 * a.c1
    k2 = a.c1.sqr() - a.c0 * a.c2
    t = ((a.c2 * k1 + a.c1 * k2).mul_nr() + a.c0 * k0).mat()     # Fq2
    n = t.c0.sqr() + t.c1.sqr()                        # Fq
    return n, (t, k0, k1, k2)


def fq12_inv_finish(f, ninv, inter):
    t, k0, k1, k2 = inter
    tinv = Fq2(t.c0 * ninv, -(t.c1 * ninv))            # fq2.go:133-147
    i6 = Fq6(k0 * tinv, k1 * tinv, k2 * tinv)          # fq6.go:330-335
    return Fq12(f.c0 * i6, -(f.c1 * i6))               # fq12.go:230-236


# ---- Miller loop (pairing.go:16-75 fused with g2.go:650-801) -----------------------------------------
BLS_X = 0xd201000000010000


def line_double(r):
    """g2.go:655-708; r = (x, y, z) Fq2 Jacobian; returns (new r, (c0, c1, c2))"""
    x, y, z = r
    t0 = x.sqr().mat()
    t1 = y.sqr().mat()
    t2 = t1.sqr().mat()
    t3 = ((t1 + x).sqr() - t0 - t2).dbl()
    t4 = t0.scale(3)
    t6 = x + t4
    t5 = t4.sqr().mat()
    zsq = z.sqr().mat()
    nx = (t5 - t3 - t3).mat()
    nz = ((z + y).sqr() - t1 - zsq).mat()
    ny = ((t3 - nx) * t4 - t2.scale(8)).mat()
    o1 = -((t4 * zsq).dbl())
    o2 = t6.sqr() - t0 - t5 - t1.scale(4)
    o0 = (nz * zsq).dbl()
    return (nx, ny, nz), (o0, o1, o2)


def line_add(r, qx, qy):
    """g2.go:710-772"""
    x, y, z = r
    zsq = z.sqr().mat()
    ysq = qy.sqr().mat()
    t0 = zsq * qx
    t1 = ((qy + z).sqr() - ysq - zsq) * zsq
    t2 = (t0 - x).mat()
    t3 = t2.sqr().mat()
    t4 = t3.scale(4)
    t5 = (t4 * t2).mat()
    t6 = (t1 - y - y).mat()
    t9 = t6 * qx
    t7 = (t4 * x).mat()
    nx = (t6.sqr() - t5 - t7 - t7).mat()
    nz = ((z + t2).sqr() - zsq - t3).mat()
    t10 = qy + nz
    t8 = (t7 - nx) * t6
    ny = (t8 - (y * t5).dbl()).mat()
    t10 = t10.sqr() - ysq - nz.sqr()
    o2 = t9.dbl() - t10
    o0 = nz.dbl()
    o1 = -(t6.dbl())
    return (nx, ny, nz), (o0, o1, o2)


def ell(f, coeffs, px, py):
    """pairing.go:28-39"""
    c0, c1, c2 = coeffs
    return f.mul_by_014(c2, c1.mul_fq(px), c0.mul_fq(py))


def miller_loop(p, pairs):
    """pairs: list of (px, py, qx, qy); one shared accumulator f (pairing.go:40-69)"""
    f = None
    rs = [(qx, qy, Fq2(p.const(1), p.const(0))) for _, _, qx, qy in pairs]
    xr = BLS_X >> 1
    def step_all(f, fn):
        for i, (px, py, qx, qy) in enumerate(pairs):
            rs[i], co = fn(rs[i], qx, qy)
            f = ell(f, co, px, py) if f is not None else ell(fq12_one(p), co, px, py)
            f = f.mat()
        return f
    for bit in range(61, -1, -1):
        f = step_all(f, lambda r, qx, qy: line_double(r))
        if (xr >> bit) & 1:
            f = step_all(f, line_add)
        f = f.sqr().mat()
    f = step_all(f, lambda r, qx, qy: line_double(r))
    return f.conj()


# ---- final exponentiation (pairing.go:79-129) ---------------------------------------------------------
def exp_by_x(f, x):
    """conj(f^x), cyclotomic squarings (pairing.go:92-98)"""
    acc = f
    top = x.bit_length() - 1
    for bit in range(top - 1, -1, -1):
        acc = acc.cyclotomic_sqr().mat()
        if (x >> bit) & 1:
            acc = (acc * f).mat()
    return acc.conj()


def final_exp(f, ninv, inter, tabs, spill=None):
    X = BLS_X
    sp = spill or (lambda v: v)
    f1 = f.conj()
    f2 = fq12_inv_finish(f, ninv, inter)
    r = (f1 * f2).mat()
    r = (r.frobenius(2, tabs) * r).mat()
    r = sp(r)
    y0 = sp(r.cyclotomic_sqr().mat())
    y1 = exp_by_x(y0, X).mat()
    y2 = exp_by_x(y1, X >> 1).mat()
    y1 = ((y1 * r.conj()).conj() * y2).mat()
    y1 = sp(y1)
    y2 = sp(exp_by_x(y1, X).mat())
    y3 = exp_by_x(y2, X).mat()
    y3 = sp((y3 * y1.conj()).mat())
    y1 = (y1.frobenius(3, tabs) * y2.frobenius(2, tabs)).mat()
    y1 = sp(y1)
    y2 = exp_by_x(y3, X).mat()
    y2 = ((y2 * y0).mat() * r).mat()
    y1 = (y1 * y2).mat()
    return (y1 * y3.frobenius(1, tabs)).mat()
