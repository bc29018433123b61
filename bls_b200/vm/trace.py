"""Symbolic tracer for the Fq2-granular virtual machine (bls_b200/csrc/vm.cuh).

The unit of data is an Fq2 element (96 bytes, c0 || c1); the unit of work is a fused operation

    dst = [xi *] (A * B  or  A^2  or  0) + sum of addends

where A and B are signed sums of up to three slot values and every addend is +/- a slot value,
optionally multiplied by xi = 1 + u (fq2.go:41-45) or conjugated.  An Fq2 product is two
two-product dot products with one Montgomery reduction each (c0 = a0*b0 + a1*(2Q - b1),
c1 = a0*b1 + a1*b0): 888 wide MACs and no Karatsuba fix-up inside Fq2.

Additions are *lazy*: an `F2` is a linear combination of slot values with small integer
coefficients and xi / conjugation flags; nothing is emitted until a combination is needed as a
multiplication operand, grows beyond LAZY_TERMS, or is an output.  A peephole pass then folds the
tower's Karatsuba recombinations into the addend list of the product that finishes last, so most
linear work rides on a multiplication instead of being a step of its own.

The formulas are the ones of bls_b200/csrc/{tower,pairing}.cuh (themselves citing the reference:
fq2.go, fq6.go, fq12.go, g2.go:655-772, pairing.go:16-129); all values are canonical residues, so
the results are bit-identical to the reference's whatever the schedule.
"""
from __future__ import annotations

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS_X = 0xd201000000010000

MAX_WEIGHT = 8           # addends per operation (sum of |coefficients|)
MAX_OPERAND = 3          # terms of a multiplication operand
LAZY_TERMS = 4           # combinations with more terms than this are materialised when created
CONST_SEG = 4            # global segment of the constant table (stride 0: shared by all units)
RELOAD_AFTER = 120       # ops after which a cached copy of a global/constant value is loaded afresh

# term key: (value id, xi flag, conj flag)


class Program:
    """ops: FMA {dst, mode: 'mul'|'sqr'|'lin', a: [term], b: [term], add: [term], xi: bool}, IO {...};
    a term is (value, sign, xi, conj)"""

    def __init__(self, name):
        self.name = name
        self.ops = []
        self.nvals = 0
        self.consts = []         # Fq2 constants as (c0, c1) canonical integers
        self._const_ids = {}
        self._lin_cache = {}
        self._glob = {}          # handle -> (seg, fq index, width in Fq: 1 or 2)
        self._glob_copy = {}
        self.spill_fq = 0

    def _new(self):
        v = self.nvals
        self.nvals += 1
        return v

    def _glob_handle(self, seg, idx, width):
        h = -(len(self._glob) + 1)
        self._glob[h] = (seg, idx, width)
        return h

    def const(self, c0, c1=0):
        key = (c0 % Q, c1 % Q)
        if key not in self._const_ids:
            self._const_ids[key] = self._glob_handle(CONST_SEG, 2 * len(self.consts), 2)
            self.consts.append(key)
        return F2(self, {(self._const_ids[key], 0, 0): 1})

    def load(self, seg, idx, width=2):
        """input element -> slot value; width 1 loads a single Fq as (x, 0)"""
        return F2(self, {(self._slot_of(self._glob_handle(seg, idx, width)), 0, 0): 1})

    def _slot_of(self, v):
        if v >= 0:
            return v
        c = self._glob_copy.get(v)
        if c is None or len(self.ops) - c[1] > RELOAD_AFTER:
            d = self._new()
            seg, idx, width = self._glob[v]
            self.ops.append({"kind": "IO", "op": "load", "dst": d, "seg": seg, "idx": idx, "width": width, "src_handle": v})
            c = (d, len(self.ops))
            self._glob_copy[v] = c
        return c[0]

    def store(self, x, seg, idx, width=2):
        m = self.materialise(x)
        h = self._glob_handle(seg, idx, width)
        self.ops.append({"kind": "IO", "op": "store", "src": m.single()[0], "seg": seg, "idx": idx, "width": width, "dst_handle": h})
        return F2(self, {(h, 0, 0): 1})

    def spill(self, x):
        r = self.store(x, 3, self.spill_fq)
        self.spill_fq += 2
        return r

    # -- emission -------------------------------------------------------------------------------
    def _terms(self, comb):
        """dict -> flat list of (slot value, sign, xi, conj), coefficient k expanded into k entries"""
        flat = []
        for (v, xi, cj), c in comb.items():
            flat += [(self._slot_of(v), 1 if c > 0 else -1, xi, cj)] * abs(c)
        return flat

    def _emit_lin(self, dst, comb):
        flat = self._terms(comb)
        if not flat:
            flat = [(self._slot_of(self.const(0).single()[0]), 1, 0, 0)]
        while len(flat) > MAX_WEIGHT:
            chunk, flat = flat[:MAX_WEIGHT], flat[MAX_WEIGHT:]
            tmp = self._new()
            self.ops.append({"kind": "FMA", "mode": "lin", "dst": tmp, "a": [], "b": [], "add": chunk, "xi": False})
            flat.append((tmp, 1, 0, 0))
        self.ops.append({"kind": "FMA", "mode": "lin", "dst": dst, "a": [], "b": [], "add": flat, "xi": False})

    def materialise(self, x):
        if x.is_single():
            return x
        if len(x.terms) == 1:
            (k, c), = x.terms.items()
            if c == 1 and k[1] == 0 and k[2] == 0:                  # a bare global / constant
                return F2(self, {(self._slot_of(k[0]), 0, 0): 1})
        key = tuple(sorted(x.terms.items()))
        if key in self._lin_cache:
            return self._lin_cache[key]
        dst = self._new()
        self._emit_lin(dst, x.terms)
        r = F2(self, {(dst, 0, 0): 1})
        self._lin_cache[key] = r
        return r

    def _operand(self, x):
        """-> list of terms (<= MAX_OPERAND, coefficients +/-1)"""
        if sum(abs(c) for c in x.terms.values()) > MAX_OPERAND:
            x = self.materialise(x)
        return self._terms(x.terms)

    def _is_const(self, x, c0, c1=0):
        if len(x.terms) != 1:
            return False
        (k, c), = x.terms.items()
        return c == 1 and k[0] < 0 and k[1] == 0 and k[2] == 0 and self._glob[k[0]][0] == CONST_SEG and \
            self.consts[self._glob[k[0]][1] // 2] == (c0 % Q, c1 % Q)

    def mul(self, a, b):
        if not a.terms or not b.terms or self._is_const(a, 0) or self._is_const(b, 0):
            return F2(self, {})
        if self._is_const(a, 1):
            return b
        if self._is_const(b, 1):
            return a
        wa, wb = sum(abs(c) for c in a.terms.values()), sum(abs(c) for c in b.terms.values())
        if min(wa, MAX_OPERAND) * min(wb, MAX_OPERAND) > 4:      # unreduced operand sums: |A| < na Q, |B| < nb Q, 2 na nb Q^2 < Q 2^384
            if wa >= wb:
                a = self.materialise(a)
            else:
                b = self.materialise(b)
        dst = self._new()
        self.ops.append({"kind": "FMA", "mode": "mul", "dst": dst, "a": self._operand(a), "b": self._operand(b), "add": [], "xi": False})
        return F2(self, {(dst, 0, 0): 1})

    def sqr(self, a):
        if not a.terms:
            return F2(self, {})
        a = self.materialise(a)                                  # (a0 + a1)(a0 - a1) needs a0, a1 < Q
        dst = self._new()
        self.ops.append({"kind": "FMA", "mode": "sqr", "dst": dst, "a": self._operand(a), "b": [], "add": [], "xi": False})
        return F2(self, {(dst, 0, 0): 1})


class F2:
    """lazy linear combination of Fq2 slot values: {(value, xi, conj): integer}"""
    __slots__ = ("p", "terms")

    def __init__(self, p, terms):
        self.p = p
        self.terms = {k: c for k, c in terms.items() if c}
        if len(self.terms) > LAZY_TERMS or sum(abs(c) for c in self.terms.values()) > 2 * MAX_WEIGHT:
            self.terms = dict(p.materialise(self).terms)

    def is_single(self):
        if len(self.terms) != 1:
            return False
        (k, c), = self.terms.items()
        return c == 1 and k[0] >= 0 and k[1] == 0 and k[2] == 0

    def single(self):
        (k, c), = self.terms.items()
        return k

    def _comb(self, o, s):
        t = dict(self.terms)
        for k, c in o.terms.items():
            t[k] = t.get(k, 0) + s * c
        return F2(self.p, t)

    def __add__(self, o): return self._comb(o, 1)
    def __sub__(self, o): return self._comb(o, -1)
    def __neg__(self): return F2(self.p, {k: -c for k, c in self.terms.items()})
    def scale(self, k): return F2(self.p, {t: k * c for t, c in self.terms.items()})
    def dbl(self): return self.scale(2)

    def mul_nr(self):                                    # * (1 + u), fq2.go:41-45
        x = self if all(k[1] == 0 for k in self.terms) else self.p.materialise(self)
        return F2(self.p, {(k[0], 1, k[2]): c for k, c in x.terms.items()})

    def conj(self):                                      # fq2.go:156-158 with power 1
        x = self if all(k[1] == 0 for k in self.terms) else self.p.materialise(self)
        return F2(self.p, {(k[0], 0, 1 - k[2]): c for k, c in x.terms.items()})

    def __mul__(self, o): return self.p.mul(self, o)
    def sqr(self): return self.p.sqr(self)
    def mat(self): return self.p.materialise(self)


# ---- tower over F2 -------------------------------------------------------------------------------------
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
        v0, v1, v2 = (a.c0 * b.c0).mat(), (a.c1 * b.c1).mat(), (a.c2 * b.c2).mat()
        x = ((a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2).mul_nr() + v0
        y = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1 + v2.mul_nr()
        z = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2 + v1
        return Fq6(x, y, z)

    def mul_by_01(self, b0, b1):                                                # fq6.go:60-90
        a = self
        v0, v1 = (a.c0 * b0).mat(), (a.c1 * b1).mat()
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
    def coeffs(self): return [self.c0, self.c1, self.c2]


class Fq12:
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def conj(self): return Fq12(self.c0, -self.c1)                              # fq12.go:27-29

    def __mul__(self, o):                                                       # fq12.go:198-213
        aa = (self.c0 * o.c0).mat()
        bb = (self.c1 * o.c1).mat()
        s = (self.c0 + self.c1) * (o.c0 + o.c1)
        return Fq12(bb.mul_nr() + aa, s - aa - bb)

    def sqr(self):                                                              # fq12.go:180-195
        a, b = self.c0, self.c1
        ab = (a * b).mat()
        s = (a + b) * (b.mul_nr() + a)
        return Fq12(s - ab - ab.mul_nr(), ab.dbl())

    def mul_by_014(self, d0, d1, d4):                                           # fq12.go:32-47
        aa = self.c0.mul_by_01(d0, d1).mat()
        bb = self.c1.mul_by_1(d4).mat()
        s = (self.c1 + self.c0).mul_by_01(d0, d1 + d4)
        return Fq12(bb.mul_nr() + aa, s - aa - bb)

    def frobenius(self, power, tabs):                                           # fq12.go:171-177
        c0 = self.c0.frobenius(power, tabs)
        c1 = self.c1.frobenius(power, tabs)
        k = tabs["fq12_c1"][power]
        return Fq12(c0, Fq6(c1.c0 * k, c1.c1 * k, c1.c2 * k))

    def cyclotomic_sqr(self):                                                   # Granger-Scott, tower.cuh
        def fp4(a, b):
            t0, t1 = a.sqr().mat(), b.sqr().mat()
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
        t3 = t3.mat().mul_nr()
        n10 = t3.scale(3) + z2.dbl()
        n02 = t2.scale(3) - z3.dbl()
        return Fq12(Fq6(n00, n01, n02), Fq6(n10, n11, n12))

    def mat(self): return Fq12(self.c0.mat(), self.c1.mat())
    def coeffs(self): return self.c0.coeffs() + self.c1.coeffs()


def fq12_from(c):
    c = list(c)
    return Fq12(Fq6(c[0], c[1], c[2]), Fq6(c[3], c[4], c[5]))


def fq12_one(p):
    z = p.const(0)
    return fq12_from([p.const(1)] + [z] * 5)


# ---- inversion down to one Fq: norm chain (fq12.go:216-237, fq6.go:295-336, fq2.go:133-147) --------
def fq12_inv_norm(f):
    """n = (t.c0^2 + t.c1^2, 0) with t the Fq2 norm of f; returns n (an F2 whose c1 is 0) and the intermediates"""
    t0 = f.c0 * f.c0 - (f.c1 * f.c1).mul_nr()          # Fq6
    a = t0.mat()
    k0 = (a.c0.sqr() - (a.c1 * a.c2).mul_nr()).mat()
    k1 = (a.c2.sqr().mul_nr() - a.c0 * a.c1).mat()
    k2 = (a.c1.sqr() - a.c0 * a.c2).mat()
    t = ((a.c2 * k1 + a.c1 * k2).mul_nr() + a.c0 * k0).mat()     # Fq2
    n = t * t.conj()                                   # (t0^2 + t1^2, 0)
    return n, (t, k0, k1, k2)


def fq12_inv_finish(f, ninv, inter):
    """ninv = (1/n, 0)"""
    t, k0, k1, k2 = inter
    tinv = (t.conj() * ninv).mat()                     # fq2.go:133-147
    i6 = Fq6((k0 * tinv).mat(), (k1 * tinv).mat(), (k2 * tinv).mat())          # fq6.go:330-335
    return Fq12(f.c0 * i6, -(f.c1 * i6))               # fq12.go:230-236


# ---- Miller loop (pairing.go:16-75 fused with g2.go:650-801) -----------------------------------------
def line_double(r):
    """g2.go:655-708; r = (x, y, z) Jacobian over Fq2; returns (new r, (c0, c1, c2))"""
    x, y, z = r
    t0 = x.sqr().mat()
    t1 = y.sqr().mat()
    t2 = t1.sqr().mat()
    t3 = ((t1 + x).sqr() - t0 - t2).dbl().mat()
    t4 = t0.scale(3).mat()
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
    t1 = ((qy + z).sqr() - ysq - zsq).mat() * zsq
    t2 = (t0 - x).mat()
    t3 = t2.sqr().mat()
    t4 = t3.scale(4).mat()
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
    """pairing.go:28-39; px, py are (x, 0) Fq2 embeddings of the G1 coordinates"""
    c0, c1, c2 = coeffs
    return f.mul_by_014(c2.mat(), (c1 * px).mat(), (c0 * py).mat())


def miller_loop(p, pairs):
    """pairs: list of (px, py, qx, qy); one shared accumulator f (pairing.go:40-69)"""
    f = None
    rs = [(qx, qy, p.const(1)) for _, _, qx, qy in pairs]
    xr = BLS_X >> 1

    def step_all(f, fn):
        for i, (px, py, qx, qy) in enumerate(pairs):
            rs[i], co = fn(rs[i], qx, qy)
            f = ell(f if f is not None else fq12_one(p), co, px, py).mat()
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
