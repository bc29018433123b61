"""Pins the CPU oracle (oracle/) against the reference's own known-answer vectors
(tests/golden/ref_kats.json, extracted by tests/golden/make_golden.py) and against Python
big-integer arithmetic, the way the reference's tests do against math/big."""
import ctypes
import random

import numpy as np

from bls_b200 import layout as L

U64 = np.uint64


def H(s):
    return int(s, 16)


def mont(v):
    return L.fp_from_int(v)


def test_constants(kats, orc):
    c = kats["constants"]
    assert H(c["q"]) == L.Q
    assert H(c["r2"]) == L.MONT_R * L.MONT_R % L.Q
    assert H(c["b_coeff_mont"]) == 4 * L.MONT_R % L.Q          # g1.go:29
    g1 = orc.g1_generator()[0]
    assert L.fp_to_int(g1["x"]) == H(c["g1_gen"][0]) and L.fp_to_int(g1["y"]) == H(c["g1_gen"][1])
    g2 = orc.g2_generator()[0]
    xc1, xc0, yc1, yc0 = [H(x) for x in c["g2_gen_xc1_xc0_yc1_yc0"]]
    assert [L.fp_to_int(g2["x"][0]), L.fp_to_int(g2["x"][1])] == [xc0, xc1]
    assert [L.fp_to_int(g2["y"][0]), L.fp_to_int(g2["y"][1])] == [yc0, yc1]
    assert orc.g1.is_on_curve(orc.g1_generator()) and orc.g2.is_on_curve(orc.g2_generator())
    assert orc.g1.in_subgroup(orc.g1_generator()) and orc.g2.in_subgroup(orc.g2_generator())


def test_limb_primitives(kats, orc):
    lib = orc.lib()
    for a, b, borrow, out, ob in kats["sub_with_borrow"]["cases"]:      # primitivefuncs_test.go:25-101
        c = ctypes.c_uint64(borrow)
        assert lib.orc_sub_with_borrow(ctypes.c_uint64(a), ctypes.c_uint64(b), ctypes.byref(c)) == out and c.value == ob
    for a, b, carry, out, oc in kats["add_with_carry"]["cases"]:        # primitivefuncs_test.go:110-186
        c = ctypes.c_uint64(carry)
        assert lib.orc_add_with_carry(ctypes.c_uint64(a), ctypes.c_uint64(b), ctypes.byref(c)) == out and c.value == oc
    for a, b, cc, carry, out, oc in kats["mac_with_carry"]["cases"]:    # primitivefuncs_test.go:188-240
        c = ctypes.c_uint64(carry)
        assert lib.orc_mac_with_carry(ctypes.c_uint64(a), ctypes.c_uint64(b), ctypes.c_uint64(cc), ctypes.byref(c)) == out
        assert c.value == oc


def _mul_repr(orc, a, b):
    a = np.array(L.int_to_limbs(a), U64); b = np.array(L.int_to_limbs(b), U64)
    hi = np.zeros(6, U64); lo = np.zeros(6, U64)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    orc.lib().orc_multiply_fq_repr(p(a), p(b), p(hi), p(lo))
    return L.limbs_to_int(hi), L.limbs_to_int(lo)


def test_multiply_fq_repr(kats, orc):
    k = kats["multiply_fq_repr"]                                          # primitivefuncs_test.go:244-262
    hi, lo = _mul_repr(orc, H(k["f0"]), H(k["f1"]))
    assert hi == H(k["hi"]) and lo == H(k["lo"])
    rng = random.Random(1)                                                # primitivefuncs_test.go:264-285 (vs big ints)
    for _ in range(2000):
        a, b = rng.getrandbits(384), rng.getrandbits(384)
        hi, lo = _mul_repr(orc, a, b)
        assert (hi << 384) | lo == a * b


def test_mont_reduce(kats, orc):
    k = kats["mont_reduce"]                                               # fqrepr_test.go:136-147
    hi = np.array(L.int_to_limbs(H(k["hi"])), U64); lo = np.array(L.int_to_limbs(H(k["lo"])), U64)
    out = np.zeros(6, U64)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    orc.lib().orc_mont_reduce(p(hi), p(lo), p(out))
    assert L.limbs_to_int(out) == H(k["expected"])


def test_fq_vs_bigint(kats, orc):
    """fq_test.go:61-166 (differential vs math/big) with the xorshift reader, plus edge values."""
    xs = orc.XorShift(1)
    a = xs.rand_fq(300); b = xs.rand_fq(300)
    edge = np.stack([mont(v) for v in (0, 1, L.Q - 1, 2, L.Q - 2, L.MONT_R, (L.Q + 1) // 2)])
    a = np.concatenate([a, edge, edge[::-1]]); b = np.concatenate([b, edge, edge])
    ai = [L.fp_to_int(x) for x in a]; bi = [L.fp_to_int(x) for x in b]
    for op, f in (("mul", lambda x, y: x * y), ("add", lambda x, y: x + y), ("sub", lambda x, y: x - y)):
        got = orc.fq(op, a, b)
        assert [L.fp_to_int(x) for x in got] == [f(x, y) % L.Q for x, y in zip(ai, bi)]
        assert all(L.limbs_to_int(x) < L.Q for x in got)
    assert [L.fp_to_int(x) for x in orc.fq("square", a)] == [x * x % L.Q for x in ai]
    assert [L.fp_to_int(x) for x in orc.fq("neg", a)] == [-x % L.Q for x in ai]
    assert [L.fp_to_int(x) for x in orc.fq("double", a)] == [2 * x % L.Q for x in ai]
    assert [L.fp_to_int(x) for x in orc.fq("inverse", a)] == [pow(x, -1, L.Q) if x else 0 for x in ai]
    v = H(kats["fq_inverse_input"]["value"])                              # fq_test.go:189-207
    assert L.fp_to_int(orc.fq("inverse", mont(v))[0]) == pow(v, -1, L.Q)
    sq = orc.fq("square", a)
    rt = orc.fq("sqrt", sq)                                               # fq_test.go:168-187
    assert [L.fp_to_int(x) ** 2 % L.Q for x in rt] == [L.fp_to_int(x) for x in sq]


def fq2_(c0, c1):
    return np.stack([mont(c0), mont(c1)])


def fq2_ints(x):
    x = np.asarray(x).reshape(2, 6)
    return [L.fp_to_int(x[0]), L.fp_to_int(x[1])]


def test_fq2_kats(kats, orc):
    k = {n: [H(v) for v in vs] for n, vs in kats["fq2"].items() if n != "cite"}
    v = k["TestFQ2Squaring"]                                              # fq2_test.go:71-98
    assert fq2_ints(orc.fq2("square", fq2_(1, 1))) == [0, 2]
    assert fq2_ints(orc.fq2("square", fq2_(0, 1))) == [L.Q - 1, 0]
    assert fq2_ints(orc.fq2("square", fq2_(v[0], v[1]))) == v[2:4]
    v = k["TestFQ2Mul"]                                                   # fq2_test.go:100-115
    assert fq2_ints(orc.fq2("mul", fq2_(v[0], v[1]), fq2_(v[2], v[3]))) == v[4:6]
    v = k["TestFQ2Inverse"]                                               # fq2_test.go:117-134
    assert fq2_ints(orc.fq2("inverse", fq2_(v[0], v[1]))) == v[2:4]
    v = k["TestFQ2Addition"]
    assert fq2_ints(orc.fq2("add", fq2_(v[0], v[1]), fq2_(v[2], v[3]))) == v[4:6]
    v = k["TestFQ2Subtraction"]
    assert fq2_ints(orc.fq2("sub", fq2_(v[0], v[1]), fq2_(v[2], v[3]))) == v[4:6]
    v = k["TestFQ2Negation"]
    assert fq2_ints(orc.fq2("neg", fq2_(v[0], v[1]))) == v[2:4]
    v = k["TestFQ2Doubling"]
    assert fq2_ints(orc.fq2("double", fq2_(v[0], v[1]))) == v[2:4]
    v = k["TestFQ2FrobeniusMap"]                                          # fq2_test.go:199-231
    a = fq2_(v[0], v[1])
    a1 = orc.fq2("frobenius", a)
    assert fq2_ints(a1) == v[4:6]
    assert fq2_ints(orc.fq2("frobenius", a1)) == v[6:8]
    v = k["TestFQ2Sqrt"]                                                  # fq2_test.go:233-246
    assert fq2_ints(orc.fq2("sqrt", fq2_(v[0], v[1]))) == v[2:4]
    assert fq2_ints(orc.fq2("sqrt", fq2_(v[4], 0))) == [0, v[5]]


def test_frobenius_tables(kats, orc):
    """The regenerated (1+u)^((q^k-1)/d) tables equal the reference's Montgomery literals."""
    f = kats["frobenius_mont"]
    c1, c2, c12 = orc.frobenius_tables()
    flat = lambda t: [L.limbs_to_int(x) for x in t.reshape(-1, 6)]
    assert flat(c1) == [H(x) for x in f["fq6_c1"]]                        # fq6.go:144-175
    assert flat(c2) == [H(x) for x in f["fq6_c2"]]                        # fq6.go:177-208
    assert flat(c12[1:]) == [H(x) for x in f["fq12_c1_from_1"]]           # fq12.go:122-168
    assert L.limbs_to_int(c12[0][0]) == L.MONT_R and L.limbs_to_int(c12[0][1]) == 0
    assert H(f["fq2_c1_1"]) == (L.Q - 1) * L.MONT_R % L.Q                  # fq2.go:149-152


def _jac(x, y, z=1):
    o = np.zeros(1, dtype=L.G1_JAC)
    o["x"][0] = mont(x); o["y"][0] = mont(y); o["z"][0] = mont(z)
    return o


def test_g1_kats(kats, orc):
    v = [H(x) for x in kats["g1_double"]["values"]]                       # g1_test.go:62-79
    aff = orc.g1.to_affine(orc.g1.double(_jac(v[0], v[1])))[0]
    assert [L.fp_to_int(aff["x"]), L.fp_to_int(aff["y"])] == v[2:4]
    v = [H(x) for x in kats["g1_add"]["values"]]                          # g1_test.go:81-104
    aff = orc.g1.to_affine(orc.g1.add(_jac(v[0], v[1]), _jac(v[2], v[3])))[0]
    assert [L.fp_to_int(aff["x"]), L.fp_to_int(aff["y"])] == v[4:6]
    # mixed addition agrees with the full addition (same inputs, z2 = 1)
    q = np.zeros(1, dtype=L.G1_AFFINE); q["x"][0] = mont(v[2]); q["y"][0] = mont(v[3])
    aff2 = orc.g1.to_affine(orc.g1.add_affine(_jac(v[0], v[1]), q))[0]
    assert aff2.tobytes() == aff.tobytes()


def test_pairing_kat(kats, orc):
    """bls.Pairing(G1One, G2One) vs RELIC, pairing_test.go:9-58: pins prepare + Miller loop + final exp."""
    exp = np.stack([mont(H(x)) for x in kats["pairing_g1_g2"]["coeffs"]]).reshape(2, 3, 2, 6)
    out = orc.pairing_batch(orc.g1_generator(), orc.g2_generator())[0]
    assert (out == exp).all()
    assert orc.g2_prepare(orc.g2_generator()).shape[0] == 68


def test_op_counts(orc):
    """Fq-multiplication counts of the reference algorithms."""
    g1, g2 = orc.g1_generator(), orc.g2_generator()
    orc.pairing_batch(g1, g2)                                              # warm the lazily built Frobenius tables
    orc.fq_mul_count(reset=True); orc.g2_prepare(g2); assert orc.fq_mul_count() == 1760
    orc.fq_mul_count(reset=True); f = orc.miller_loop(g1, g2); assert orc.fq_mul_count() == 1760 + 5156
    orc.fq_mul_count(reset=True); orc.final_exp(f); assert orc.fq_mul_count() == 19630


def test_bilinearity_and_compare(orc):
    """e(aP, bQ) == e(P, Q)^(ab) and CompareTwoPairings (pairing.go:140-147), incl. a negative case."""
    xs = orc.XorShift(7)
    a, b = xs.rand_fr(1), xs.rand_fr(1)
    ai, bi = L.scalar_to_int(a[0]), L.scalar_to_int(b[0])
    g1, g2 = orc.g1_generator(), orc.g2_generator()
    aP = orc.g1.mul_fr(g1, a); bQ = orc.g2.mul_fr(g2, b)
    ab = L.scalar_from_int(ai * bi % L.R_ORDER).reshape(1, 4)
    abP = orc.g1.mul_fr(g1, ab)
    assert orc.compare_two_pairings(aP, bQ, abP, orc.g2.to_proj(g2))
    assert not orc.compare_two_pairings(aP, bQ, orc.g1.to_proj(g1), orc.g2.to_proj(g2))
    lhs = orc.pairing_batch(orc.g1.to_affine(aP), orc.g2.to_affine(bQ))[0]
    rhs = orc.pairing_batch(orc.g1.to_affine(abP), g2)[0]
    assert (lhs == rhs).all()
    # product form: e(aP, bQ) * e(-abP, Q) == 1
    nabP = orc.g1.to_affine(abP).copy()
    nabP["y"][0] = orc.fq("neg", nabP["y"][0])[0]
    p = np.concatenate([orc.g1.to_affine(aP), nabP]); q = np.concatenate([orc.g2.to_affine(bQ), g2])
    assert orc.pairing_product_is_one(p, q, [0, 2]).tolist() == [1]
    assert orc.pairing_product_is_one(p, q, [0, 1, 2]).tolist() == [0, 0]


def test_compress_roundtrip(orc):
    """CompressG1/G2 <-> DecompressG1/G2 (g1.go:185-249, g2.go:219-289); g1pubs/bls_test.go:344-384 style."""
    xs = orc.XorShift(5)
    s = xs.rand_fr(3)
    p = orc.g1.to_affine(orc.g1.mul_fr(np.repeat(orc.g1_generator(), 3), s))
    q = orc.g2.to_affine(orc.g2.mul_fr(np.repeat(orc.g2_generator(), 3), s))
    for i in range(3):
        b = orc.g1.compress(p[i:i + 1]); assert len(b) == 48 and b[0] & 0x80
        e, back = orc.g1.decompress(b); assert e == 0 and back.tobytes() == p[i:i + 1].tobytes()
        b = orc.g2.compress(q[i:i + 1]); assert len(b) == 96 and b[0] & 0x80
        e, back = orc.g2.decompress(b); assert e == 0 and back.tobytes() == q[i:i + 1].tobytes()
    inf = np.zeros(1, dtype=L.G1_AFFINE); inf["inf"] = 1
    b = orc.g1.compress(inf); assert b[0] == 0xC0 and not any(b[1:])
    e, back = orc.g1.decompress(b); assert e == 0 and back["inf"][0] == 1
    assert orc.g1.decompress(bytes(48))[0] == 1                            # "unexpected compression mode"
    assert orc.g1.decompress(bytes([0xC0] + [0] * 46 + [1]))[0] == 2       # junk in compressed infinity
