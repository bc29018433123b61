#!/usr/bin/env python3
"""Generate bls_b200/csrc/fp_mul_asm.inc: the 384-bit Montgomery multiplication and squaring of
the BLS12-381 base field as single inline-PTX blocks over 12 x u32 limbs.

Replaces MultiplyFQRepr + MontReduce (stub_fallback.go:11-116, primitivefuncs_amd64.s:79-1467):
same function (a*b*2^-384 mod Q), different schedule -- the 32x32->64 products of one operand word
are split into an "even" and an "odd" accumulator so that every mad.lo.cc/madc.hi.cc pair lands on
an aligned register pair (one IMAD.WIDE.U32 with carry in SASS) and the two accumulators give the
scheduler two independent carry chains.  Reduction is interleaved word by word.

The same instruction list is executed by a Python interpreter below (`--selftest`) and compared
with big-integer arithmetic, so the schedule is verified without a GPU.
"""
import random
import sys
import pathlib

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
N = 12
W = 1 << 32
MODL = [(Q >> (32 * i)) & (W - 1) for i in range(N)]
M0 = (-pow(Q, -1, W)) % W          # 0xfffcfffd, low word of stub_fallback.go:59


class Prog:
    """tiny PTX-subset builder: registers are names; immediates are ints"""

    def __init__(self):
        self.ins = []
        self.ntmp = 0

    def tmp(self):
        self.ntmp += 1
        return "t%d" % (self.ntmp - 1)

    def emit(self, op, d, *src):
        self.ins.append((op, d) + src)


def mul_row(p, acc, a, a0, bi):
    """acc[j], acc[j+1] = a[a0+j] * bi for even j (no carries: disjoint pairs)"""
    for j in range(0, N, 2):
        p.emit("mul.lo", acc[j], a[a0 + j], bi)
        p.emit("mul.hi", acc[j + 1], a[a0 + j], bi)


def cmad_row(p, acc, a, a0, bi):
    """acc += sum_j a[a0+j]*bi*W^j over even j, one carry chain; leaves the carry-out in CC"""
    for j in range(0, N, 2):
        p.emit("mad.lo.cc" if j == 0 else "madc.lo.cc", acc[j], bi, a[a0 + j], acc[j])
        p.emit("madc.hi.cc", acc[j + 1], bi, a[a0 + j], acc[j + 1])


def madc_row_rshift(p, odd, a, a0, bi):
    """odd = (odd >> 64) + sum_j a[a0+j]*bi*W^j, carry-in from CC"""
    for j in range(0, N - 2, 2):
        p.emit("madc.lo.cc", odd[j], a[a0 + j], bi, odd[j + 2])
        p.emit("madc.hi.cc", odd[j + 1], a[a0 + j], bi, odd[j + 3])
    p.emit("madc.lo.cc", odd[N - 2], a[a0 + N - 2], bi, 0)
    p.emit("madc.hi", odd[N - 1], a[a0 + N - 2], bi, 0)


def mad_redc(p, even, odd, a, bi, first):
    if first:
        mul_row(p, odd, a, 1, bi)
        mul_row(p, even, a, 0, bi)
    else:
        p.emit("add.cc", even[0], even[0], odd[1])
        madc_row_rshift(p, odd, a, 1, bi)
        cmad_row(p, even, a, 0, bi)
        p.emit("addc", odd[N - 1], odd[N - 1], 0)
    mi = p.tmp()
    p.emit("mul.lo", mi, even[0], M0)
    cmad_row(p, odd, MODL, 1, mi)
    cmad_row(p, even, MODL, 0, mi)
    p.emit("addc", odd[N - 1], odd[N - 1], 0)


def final_sub(p, r, x):
    """r = x - Q if x >= Q else x   (x < 2Q)"""
    d = [p.tmp() for _ in range(N)]
    for i in range(N):
        p.emit("sub.cc" if i == 0 else "subc.cc", d[i], x[i], MODL[i])
    bw = p.tmp()
    p.emit("subc", bw, 0, 0)            # 0 or 0xffffffff (borrow)
    for i in range(N):
        p.emit("selb", r[i], x[i], d[i], bw)   # r = bw ? x : d


def finish(p, even, reduce):
    if reduce:
        final_sub(p, ["r%d" % i for i in range(N)], even)
    else:                                   # lazily reduced result in [0, 2Q): the caller reduces after adding its addends
        for i in range(N):
            p.emit("mov", "r%d" % i, even[i])


def gen_mul(reduce=True):
    p = Prog()
    a = ["a%d" % i for i in range(N)] + [0]
    b = ["b%d" % i for i in range(N)]
    even = [p.tmp() for _ in range(N)]
    odd = [p.tmp() for _ in range(N)]
    for i in range(0, N, 2):
        mad_redc(p, even, odd, a, b[i], i == 0)
        mad_redc(p, odd, even, a, b[i + 1], False)
    p.emit("add.cc", even[0], even[0], odd[1])
    for i in range(1, N - 1):
        p.emit("addc.cc", even[i], even[i], odd[i + 1])
    p.emit("addc", even[N - 1], even[N - 1], 0)
    finish(p, even, reduce)
    return p


def mad2_redc(p, even, odd, a, bi, c, di, first):
    """one word of the two-product accumulation: acc += a*bi + c*di, then one reduction word"""
    if first:
        mul_row(p, odd, a, 1, bi)
        mul_row(p, even, a, 0, bi)
    else:
        p.emit("add.cc", even[0], even[0], odd[1])
        madc_row_rshift(p, odd, a, 1, bi)
        cmad_row(p, even, a, 0, bi)
        p.emit("addc", odd[N - 1], odd[N - 1], 0)
    cmad_row(p, odd, c, 1, di)
    cmad_row(p, even, c, 0, di)
    p.emit("addc", odd[N - 1], odd[N - 1], 0)
    mi = p.tmp()
    p.emit("mul.lo", mi, even[0], M0)
    cmad_row(p, odd, MODL, 1, mi)
    cmad_row(p, even, MODL, 0, mi)
    p.emit("addc", odd[N - 1], odd[N - 1], 0)


def gen_dot2(reduce=True):
    """r = (a*b + c*d) * 2^-384 mod Q with ONE interleaved reduction (2*144 + 156 wide MACs):
    the lazily reduced Fq2 product rows c0 = a0*b0 + a1*(Q - b1), c1 = a0*b1 + a1*b0.
    Operands may be as large as 2Q (sums formed on the fly): a*b + c*d < 8Q^2 < Q*2^384."""
    p = Prog()
    a = ["a%d" % i for i in range(N)] + [0]
    b = ["b%d" % i for i in range(N)]
    c = ["c%d" % i for i in range(N)] + [0]
    d = ["d%d" % i for i in range(N)]
    even = [p.tmp() for _ in range(N)]
    odd = [p.tmp() for _ in range(N)]
    for i in range(0, N, 2):
        mad2_redc(p, even, odd, a, b[i], c, d[i], i == 0)
        mad2_redc(p, odd, even, a, b[i + 1], c, d[i + 1], False)
    p.emit("add.cc", even[0], even[0], odd[1])
    for i in range(1, N - 1):
        p.emit("addc.cc", even[i], even[i], odd[i + 1])
    p.emit("addc", even[N - 1], even[N - 1], 0)
    finish(p, even, reduce)
    return p


def redc_only(p, even, odd, first):
    """one reduction word on the even/odd pair without adding a product row"""
    if not first:
        p.emit("add.cc", even[0], even[0], odd[1])
        for j in range(0, N - 2):
            p.emit("addc.cc", odd[j], odd[j + 2], 0)
        p.emit("addc.cc", odd[N - 2], 0, 0)
        p.emit("addc", odd[N - 1], 0, 0)
    mi = p.tmp()
    p.emit("mul.lo", mi, even[0], M0)
    cmad_row(p, odd, MODL, 1, mi)
    cmad_row(p, even, MODL, 0, mi)
    p.emit("addc", odd[N - 1], odd[N - 1], 0)


def gen_sqr():
    """Dedicated squaring: off-diagonal products once, doubled, plus the diagonal -> 24-limb
    square, then 12 reduction words (same quotient/carry structure as gen_mul)."""
    p = Prog()
    a = ["a%d" % i for i in range(N)]
    # 24-limb product t = a*a.  Off-diagonal part: sum_{i<j} a[i]a[j] W^(i+j), via even/odd
    # accumulators over product position parity.
    ev = [p.tmp() for _ in range(2 * N)]   # positions 0..23
    od = [p.tmp() for _ in range(2 * N)]   # positions 1..24 (od[k] is position k+1)
    used_ev = [False] * (2 * N)
    used_od = [False] * (2 * N)

    def mac_chain(acc, used, base_fn, i):
        """row i: products a[i]*a[j], j>i with (i+j) parity fixed, as one carry chain."""
        pass

    # straightforward formulation: for each i, the row of products a[i]*a[j] for j>i.  Products with
    # even (i+j) go to ev at position i+j; odd (i+j) go to od at index i+j-1.
    for i in range(N - 1):
        for par, acc, used in ((0, ev, used_ev), (1, od, used_od)):
            js = [j for j in range(i + 1, N) if (i + j) % 2 == par]
            if not js:
                continue
            first = True
            last_idx = None
            for j in js:
                idx = (i + j) - par        # index into acc of the low word
                for half, k in (("lo", idx), ("hi", idx + 1)):
                    if used[k]:
                        op = ("mad.%s.cc" if first else "madc.%s.cc") % half
                        p.emit(op, acc[k], a[i], a[j], acc[k])
                    else:
                        # fresh word: still take the carry-in of the chain
                        if first:
                            p.emit("mul.%s" % half, acc[k], a[i], a[j])
                            # a mul does not define CC; emulate with mad + 0 to start the chain cleanly
                            p.ins[-1] = ("mad.%s.cc" % half, acc[k], a[i], a[j], 0)
                        else:
                            p.emit("madc.%s.cc" % half, acc[k], a[i], a[j], 0)
                        used[k] = True
                    first = False
                last_idx = idx + 1
            # propagate the chain's carry-out into the following words of this accumulator
            k = last_idx + 1
            while k < 2 * N:
                if used[k]:
                    p.emit("addc.cc", acc[k], acc[k], 0)
                else:
                    p.emit("addc.cc", acc[k], 0, 0)
                    used[k] = True
                k += 1
                # the running sum is < W^24, a carry can only travel a few words; stop after 2
                if k > last_idx + 2:
                    break
    for k in range(2 * N):
        if not used_ev[k]:
            p.emit("mov", ev[k], 0)
        if not used_od[k]:
            p.emit("mov", od[k], 0)
    # t = ev + W*od  (positions 0..23)
    t = [p.tmp() for _ in range(2 * N)]
    p.emit("mov", t[0], ev[0])
    p.emit("add.cc", t[1], ev[1], od[0])
    for k in range(2, 2 * N):
        p.emit("addc.cc" if k < 2 * N - 1 else "addc", t[k], ev[k], od[k - 1])
    # double
    p.emit("add.cc", t[0], t[0], t[0])
    for k in range(1, 2 * N):
        p.emit("addc.cc" if k < 2 * N - 1 else "addc", t[k], t[k], t[k])
    # add the diagonal a[i]^2 at position 2i
    for i in range(N):
        p.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", t[2 * i], a[i], a[i], t[2 * i])
        p.emit("madc.hi.cc" if i < N - 1 else "madc.hi", t[2 * i + 1], a[i], a[i], t[2 * i + 1])
    # Montgomery reduction of the 24-limb t, word by word, with even/odd rows of Q*mi
    # acc window: even = positions i..i+11 (+carry), odd = positions i+1..i+12
    even = [p.tmp() for _ in range(N)]
    odd = [p.tmp() for _ in range(N)]
    for k in range(N):
        p.emit("mov", even[k], t[k])
        p.emit("mov", odd[k], 0)
    cur_e, cur_o = even, odd
    for i in range(N):
        if i > 0:
            # shift: new even = old odd (+ old even[1] into [0]); new odd = old even >> 64; bring in t[i+11]
            ne, no = cur_o, cur_e
            p.emit("add.cc", ne[0], ne[0], no[1])
            for j in range(0, N - 2):
                p.emit("addc.cc", no[j], no[j + 2], 0)
            p.emit("addc.cc", no[N - 2], 0, 0)
            p.emit("addc", no[N - 1], 0, 0)
            # incoming high limb of t at new position 11 -> index 10 of the odd accumulator
            p.emit("add.cc", no[N - 2], no[N - 2], t[i + N - 1])
            p.emit("addc", no[N - 1], no[N - 1], 0)
            cur_e, cur_o = ne, no
        mi = p.tmp()
        p.emit("mul.lo", mi, cur_e[0], M0)
        cmad_row(p, cur_o, MODL, 1, mi)
        cmad_row(p, cur_e, MODL, 0, mi)
        p.emit("addc", cur_o[N - 1], cur_o[N - 1], 0)
    # result = (cur_o positions 1..12) + (cur_e[1..11] -> positions 0..10) + t[23] at position 11
    res = cur_o
    p.emit("add.cc", res[0], res[0], cur_e[1])
    for k in range(1, N - 1):
        p.emit("addc.cc", res[k], res[k], cur_e[k + 1])
    p.emit("addc", res[N - 1], res[N - 1], t[2 * N - 1])
    final_sub(p, ["r%d" % i for i in range(N)], res)
    return p


# ---------------------------------------------------------------------------
# interpreter
# ---------------------------------------------------------------------------
def run(p, regs):
    cc = 0
    regs = dict(regs)

    def val(x):
        return x if isinstance(x, int) else regs[x]

    for ins in p.ins:
        op, d = ins[0], ins[1]
        s = [val(x) for x in ins[2:]]
        if op == "mul.lo":
            regs[d] = (s[0] * s[1]) % W
        elif op == "mul.hi":
            regs[d] = (s[0] * s[1]) >> 32
        elif op in ("mad.lo.cc", "madc.lo.cc", "mad.hi.cc", "madc.hi.cc", "madc.hi", "madc.lo"):
            prod = s[0] * s[1]
            part = prod % W if ".lo" in op else prod >> 32
            t = part + s[2] + (cc if op.startswith("madc") else 0)
            regs[d] = t % W
            if op.endswith(".cc"):
                cc = t >> 32
        elif op in ("add.cc", "addc.cc", "addc"):
            t = s[0] + s[1] + (cc if op.startswith("addc") else 0)
            regs[d] = t % W
            if op.endswith(".cc"):
                cc = t >> 32
        elif op in ("sub.cc", "subc.cc", "subc"):
            t = s[0] - s[1] - (cc if op.startswith("subc") else 0)
            regs[d] = t % W
            if op.endswith(".cc"):
                cc = 1 if t < 0 else 0
        elif op == "selb":
            regs[d] = s[0] if s[2] else s[1]
        elif op == "mov":
            regs[d] = s[0]
        else:
            raise ValueError(op)
        assert 0 <= regs[d] < W
    return regs


def limbs(v):
    return [(v >> (32 * i)) & (W - 1) for i in range(N)]


def selftest():
    rng = random.Random(20261017)
    Rinv = pow(1 << 384, -1, Q)
    pm, ps = gen_mul(), gen_sqr()
    edge = [0, 1, Q - 1, Q - 2, (1 << 384) % Q, (1 << 380), Q >> 1]
    cases = [(x, y) for x in edge for y in edge] + [(rng.randrange(Q), rng.randrange(Q)) for _ in range(3000)]
    for x, y in cases:
        regs = {"a%d" % i: v for i, v in enumerate(limbs(x))}
        regs.update({"b%d" % i: v for i, v in enumerate(limbs(y))})
        out = run(pm, regs)
        got = sum(out["r%d" % i] << (32 * i) for i in range(N))
        assert got == x * y * Rinv % Q, (hex(x), hex(y))
        out = run(ps, {"a%d" % i: v for i, v in enumerate(limbs(x))})
        got = sum(out["r%d" % i] << (32 * i) for i in range(N))
        assert got == x * x * Rinv % Q, hex(x)
    # two-product form with operands up to 2Q (and the plain multiplication on the same range)
    pd = gen_dot2()
    big = [0, 1, Q - 1, Q, 2 * Q - 1, 2 * Q - 2, Q + 1]
    cases2 = [(w, x, y, z) for w in big for x in big for y in big[:4] for z in big[3:]]
    cases2 += [tuple(rng.randrange(2 * Q) for _ in range(4)) for _ in range(3000)]
    for w, x, y, z in cases2:
        regs = {}
        for nm_, v in (("a", w), ("b", x), ("c", y), ("d", z)):
            regs.update({"%s%d" % (nm_, i): lv for i, lv in enumerate(limbs(v))})
        out = run(pd, regs)
        got = sum(out["r%d" % i] << (32 * i) for i in range(N))
        assert got == (w * x + y * z) * Rinv % Q, (hex(w), hex(x), hex(y), hex(z))
        out = run(pm, {k: v for k, v in regs.items() if k[0] in "ab"})
        got = sum(out["r%d" % i] << (32 * i) for i in range(N))
        assert got == w * x * Rinv % Q
    # lazily reduced variants: same residue, value below 2Q
    pmn, pdn = gen_mul(False), gen_dot2(False)
    for w, x, y, z in cases2[::3]:
        regs = {}
        for nm_, v in (("a", w), ("b", x), ("c", y), ("d", z)):
            regs.update({"%s%d" % (nm_, i): lv for i, lv in enumerate(limbs(v))})
        got = sum(run(pdn, regs)["r%d" % i] << (32 * i) for i in range(N))
        assert got < 2 * Q and got % Q == (w * x + y * z) * Rinv % Q
        got = sum(run(pmn, {k: v for k, v in regs.items() if k[0] in "ab"})["r%d" % i] << (32 * i) for i in range(N))
        assert got < 2 * Q and got % Q == w * x * Rinv % Q
    nd = sum(1 for i in pd.ins if i[0].startswith(("mul", "mad")))
    print("dot2 ok: %d cases, %d ins (%d mul/mad)" % (len(cases2), len(pd.ins), nd))
    nm = sum(1 for i in pm.ins if i[0].startswith(("mul", "mad")))
    ns = sum(1 for i in ps.ins if i[0].startswith(("mul", "mad")))
    print("selftest ok: %d cases; mul: %d ins (%d mul/mad), sqr: %d ins (%d mul/mad)" %
          (len(cases), len(pm.ins), nm, len(ps.ins), ns))


# ---------------------------------------------------------------------------
# PTX printer
# ---------------------------------------------------------------------------
def to_ptx(p, name, nin):
    """C macro body: asm block with operands r0..r11 (out), a0..a11, [b0..b11] (in)"""
    opn = {}
    k = 0
    for i in range(N):
        opn["r%d" % i] = "%%%d" % k; k += 1
    for i in range(N):
        opn["a%d" % i] = "%%%d" % k; k += 1
    if nin >= 2:
        for i in range(N):
            opn["b%d" % i] = "%%%d" % k; k += 1
    if nin == 4:
        for nm_ in "cd":
            for i in range(N):
                opn["%s%d" % (nm_, i)] = "%%%d" % k; k += 1

    def o(x):
        if isinstance(x, int):
            return "0x%08x" % x
        return opn.get(x, x)

    lines = ["{", ".reg .u32 t<%d>;" % p.ntmp]
    for ins in p.ins:
        op, d = ins[0], ins[1]
        s = ins[2:]
        if op == "selb":
            # r = bw ? x : d   ->  slct needs a signed compare; use setp + selp via a predicate-free trick:
            lines.append("{ .reg .pred q; setp.ne.u32 q, %s, 0; selp.u32 %s, %s, %s, q; }" % (o(s[2]), o(d), o(s[0]), o(s[1])))
        elif op == "mov":
            lines.append("mov.u32 %s, %s;" % (o(d), o(s[0])))
        else:
            lines.append("%s.u32 %s, %s;" % (op, o(d), ", ".join(o(x) for x in s)))
    lines.append("}")
    body = " \\\n".join('    "%s\\n\\t"' % l for l in lines)
    return "#define %s \\\n%s\n" % (name, body)


def main():
    if "--selftest" in sys.argv:
        selftest()
        return
    selftest()
    out = pathlib.Path(__file__).resolve().parent.parent / "bls_b200" / "csrc" / "fp_mul_asm.inc"
    txt = "// GENERATED by tools/gen_fp_asm.py -- do not edit.\n"
    txt += "// FP_MUL_PTX: operands %0..%11 = r (out), %12..%23 = a, %24..%35 = b\n"
    txt += to_ptx(gen_mul(), "FP_MUL_PTX", 2)
    txt += "// FP_DOT2_PTX: %0..%11 = r (out), %12..%23 = a, %24..%35 = b, %36..%47 = c, %48..%59 = d;  r = (a*b + c*d) * 2^-384 mod Q\n"
    txt += to_ptx(gen_dot2(), "FP_DOT2_PTX", 4)
    txt += "// FP_SQR_PTX: %0..%11 = r (out), %12..%23 = a;  r = a*a * 2^-384 mod Q -- FQ.SquareAssign (fq.go:151-198): 66 cross products once, doubled, 12 diagonal products\n"
    txt += to_ptx(gen_sqr(), "FP_SQR_PTX", 1)
    txt += "// FP_MUL_NR_PTX / FP_DOT2_NR_PTX: the same without the final conditional subtraction: result in [0, 2Q)\n"
    txt += to_ptx(gen_mul(False), "FP_MUL_NR_PTX", 2)
    txt += to_ptx(gen_dot2(False), "FP_DOT2_NR_PTX", 4)
    out.write_text(txt)
    print("wrote", out)


if __name__ == "__main__":
    main()
