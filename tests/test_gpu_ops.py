"""Operation-level parity of the DEVICE build against the oracle and the reference's own known answers, through the
b381_test_op hook of the C ABI: the PTX Montgomery multiplier, the add / sub / negate carry chains, the integer inversion, the
Fq2 / Fq6 / Fq12 tower in both the one-element-per-thread form (csrc/tower.cuh, pairing.cuh) and the four-lanes-per-element
form (csrc/quad.cuh), and the G1 / G2 group law (csrc/curve.cuh).  Reference vectors: fq_test.go:189-207 (inverse),
fq2_test.go:71-246, g1_test.go:62-104; edge operands 0, 1, Q - 1, R mod Q; xorshift-random operands as the reference's
property tests use (fq2_test.go:302-481, fq6_test.go:93-272, fq12_test.go:74-253).  Bit-exact."""
import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L

pytestmark = pytest.mark.gpu
U64 = np.uint64
H = lambda s: int(s, 16)


@pytest.fixture(scope="module")
def ctx():
    from bls_b200 import capi
    c = capi.Ctx(0)
    yield c
    c.close()


def mont(v):
    return L.fp_from_int(v)


def _edge_fq(orc, seed, n):
    """n random canonical residues with the edge operands in front: 0, 1, Q - 1, R mod Q (the limb integer 1), 2^383-ish"""
    a = orc.XorShift(seed).rand_fq(n)
    raw = lambda v: np.array(L.int_to_limbs(v), U64)
    edges = [raw(0), mont(1), raw(L.Q - 1), raw(1), raw(L.Q - 2), mont(L.Q - 1), raw((1 << 380) + 12345)]
    for i, e in enumerate(edges):
        a[i] = e
    return a


def test_fq_ops_edges_and_random(ctx, orc):
    n = 1 << 12
    a = _edge_fq(orc, 1001, n); b = np.roll(_edge_fq(orc, 1002, n), 3, axis=0)
    for op, name in [(0, "mul"), (1, "add"), (2, "sub"), (3, "square"), (4, "neg"), (5, "double"), (10, "mul"), (11, "square")]:
        out, _, _ = ctx.test_op(0, op, a, b)
        assert (out == orc.fq(name, a, b)).all(), name
    # every pairing of the edge operands with each other (7 x 7)
    ea = np.repeat(a[:7], 7, axis=0); eb = np.tile(a[:7], (7, 1))
    for op, name in [(0, "mul"), (1, "add"), (2, "sub"), (11, "square")]:
        out, _, _ = ctx.test_op(0, op, ea, eb)
        assert (out == orc.fq(name, ea, eb)).all(), name
    # the two-product dot product with one reduction (the rows of every Fq2 product): a b + b a = 2 a b
    out, _, _ = ctx.test_op(0, 9, a, b)
    assert (out == orc.fq("double", orc.fq("mul", a, b))).all()


def test_fq_mul_2_20_random_pairs(ctx, orc):
    """the size SURVEY.md section 7 step 3 asks for; the oracle multiplies 2^20 pairs in about a second"""
    n = 1 << 20
    rng = np.random.RandomState(20)
    def rnd():
        x = rng.randint(0, 1 << 32, size=(n, 12), dtype=np.uint64)
        v = (x[:, 0::2] | (x[:, 1::2] << np.uint64(32))).astype(U64)
        v[:, 5] &= np.uint64((1 << 60) - 1)            # below 2^380 < Q: canonical
        return v
    a, b = rnd(), rnd()
    out, _, _ = ctx.test_op(0, 0, a, b)
    assert (out == orc.fq("mul", a, b)).all()


def test_fq_inverse_kat_and_random(ctx, orc, kats):
    v = H(kats["fq_inverse_input"]["value"])                              # fq_test.go:189-207
    a = _edge_fq(orc, 1003, 256)
    a[7] = mont(v)
    for op in (6, 8):                                                     # almost-inverse on the limbs, Fermat chain
        out, _, _ = ctx.test_op(0, op, a)
        assert L.fp_to_int(out[7]) == pow(v, -1, L.Q)
        want = orc.fq("inverse", a)
        nz = [i for i in range(256) if L.limbs_to_int(a[i]) != 0]
        assert (out[nz] == want[nz]).all()
        assert L.limbs_to_int(out[0]) == 0                                # 0 -> 0 (the reference reports "no inverse")


def fq2_(c0, c1):
    return np.stack([mont(c0), mont(c1)])[None]


def fq2_ints(x):
    x = np.asarray(x).reshape(2, 6)
    return [L.fp_to_int(x[0]), L.fp_to_int(x[1])]


def test_fq2_reference_kats_on_device(ctx, kats):
    """fq2_test.go:71-246 against the device Fq2 routines (family 1)"""
    k = {n: [H(v) for v in vs] for n, vs in kats["fq2"].items() if n != "cite"}
    f2 = lambda op, a, b=None: fq2_ints(ctx.test_op(1, op, a, b)[0][0])
    v = k["TestFQ2Squaring"]
    assert f2(3, fq2_(1, 1)) == [0, 2]
    assert f2(3, fq2_(0, 1)) == [L.Q - 1, 0]
    assert f2(3, fq2_(v[0], v[1])) == v[2:4]
    v = k["TestFQ2Mul"]
    assert f2(0, fq2_(v[0], v[1]), fq2_(v[2], v[3])) == v[4:6]
    v = k["TestFQ2Inverse"]
    assert f2(6, fq2_(v[0], v[1])) == v[2:4]
    v = k["TestFQ2Addition"]
    assert f2(1, fq2_(v[0], v[1]), fq2_(v[2], v[3])) == v[4:6]
    v = k["TestFQ2Subtraction"]
    assert f2(2, fq2_(v[0], v[1]), fq2_(v[2], v[3])) == v[4:6]
    v = k["TestFQ2Negation"]
    assert f2(4, fq2_(v[0], v[1])) == v[2:4]
    v = k["TestFQ2Doubling"]
    assert f2(5, fq2_(v[0], v[1])) == v[2:4]
    v = k["TestFQ2FrobeniusMap"]                                          # the map is conjugation (fq2.go:156-158)
    assert f2(12, fq2_(v[0], v[1])) == v[4:6]
    assert f2(12, fq2_(v[4], v[5])) == v[6:8]


def test_fq2_fq6_random_vs_oracle(ctx, orc):
    n = 512
    a = _edge_fq(orc, 2001, 2 * n).reshape(n, 2, 6); b = orc.XorShift(2002).rand_fq(2 * n).reshape(n, 2, 6)
    for op, name in [(0, "mul"), (1, "add"), (2, "sub"), (3, "square"), (4, "neg"), (5, "double"), (6, "inverse"),
                     (10, "mul_by_nonresidue"), (12, "frobenius")]:
        out, _, _ = ctx.test_op(1, op, a, b)
        want = orc.fq2(name, a, b).reshape(n, 12)
        if name == "inverse":
            nz = [i for i in range(n) if a[i].any()]
            assert (out[nz] == want[nz]).all() and not out[[i for i in range(n) if not a[i].any()]].any()
        else:
            assert (out == want).all(), name
    a6 = _edge_fq(orc, 2003, 6 * n).reshape(n, 3, 2, 6); b6 = orc.XorShift(2004).rand_fq(6 * n).reshape(n, 3, 2, 6)
    a6[9] = 0; a6[10] = 0; a6[10, 0, 0] = mont(1)
    for op, name, arg in [(0, "mul", 0), (1, "add", 0), (2, "sub", 0), (4, "neg", 0), (7, "frobenius", 1), (7, "frobenius", 2),
                          (7, "frobenius", 3), (10, "mul_by_nonresidue", 0), (13, "mul_by_01", 0), (14, "mul_by_1", 0)]:
        out, _, _ = ctx.test_op(2, op, a6, b6, arg)
        assert (out == orc.fq6(name, a6, b6, arg).reshape(n, 36)).all(), (name, arg)
    nz = [i for i in range(n) if i != 9]
    out, _, _ = ctx.test_op(2, 6, a6)
    assert (out[nz] == orc.fq6("inverse", a6).reshape(n, 36)[nz]).all()


@pytest.mark.parametrize("family", [3, 4, 7], ids=["thread", "quad", "duo"])
def test_fq12_ops_vs_oracle(ctx, orc, family):
    """both device forms of the Fq12 routines: one element per thread and four lanes per element"""
    n = 67                                                                # ragged: the last warp of the quad form is partial
    a = orc.XorShift(3001).rand_fq(12 * n).reshape(n, 2, 3, 2, 6); b = orc.XorShift(3002).rand_fq(12 * n).reshape(n, 2, 3, 2, 6)
    a[1] = 0; a[1, 0, 0, 0] = mont(1)
    a[2, 1] = 0
    a[3] = np.array(L.int_to_limbs(L.Q - 1), U64)
    a[4] = 0
    for op, name, arg in [(0, "mul", 0), (3, "square", 0), (7, "frobenius", 1), (7, "frobenius", 2), (7, "frobenius", 3),
                          (12, "conjugate", 0), (13, "mul_by_014", 0)]:
        out, extra, _ = ctx.test_op(family, op, a, b, arg)
        assert (out == orc.fq12(name, a, b, arg).reshape(n, 72)).all(), (name, arg)
        if family == 4 and op == 13:
            assert (extra.reshape(n, 2, 6) == orc.fq2("mul", b[:, 1, 0], b[:, 1, 2])).all()
    out, _, ok = ctx.test_op(family, 6, a)
    assert ok[4] == 0 and ok[[i for i in range(n) if i != 4]].all()
    nz = [i for i in range(n) if i != 4]
    assert (out[nz] == orc.fq12("inverse", a).reshape(n, 72)[nz]).all()


@pytest.mark.parametrize("family", [3, 4, 7], ids=["thread", "quad", "duo"])
def test_cyclotomic_ops_vs_oracle(ctx, orc, family):
    n = 11
    P = hg.g1_progression(3, 1, n); Q = hg.g2_progression(4, 1, n)
    f = orc.pairing_batch(P, Q).view(U64).reshape(n, 2, 3, 2, 6).copy()
    f[5] = 0; f[5, 0, 0, 0] = mont(1)                                     # the degenerate value 1 among ordinary ones (same warp)
    out, _, _ = ctx.test_op(family, 15, f)
    assert (out == orc.fq12("square", f).reshape(n, 72)).all()
    for x in (L.BLS_X, L.BLS_X >> 1):
        want = orc.fq12("conjugate", orc.fq12("exp", f, None, x)).reshape(n, 72)
        for op in (16, 17):
            out, _, _ = ctx.test_op(family, op, f, None, x)
            assert (out == want).all(), (op, hex(x))


def _aff_from_ints(dt, x, y):
    o = np.zeros(1, dtype=dt)
    o["x"][0] = mont(x); o["y"][0] = mont(y)
    return o


def test_g1_reference_kats_on_device(ctx, orc, kats):
    """g1_test.go:62-104 (doubling, addition) against the device group law (XYZZ and Jacobian forms)"""
    v = [H(x) for x in kats["g1_double"]["values"]]
    p = _aff_from_ints(L.G1_AFFINE, v[0], v[1])
    for op in (1, 4):
        aff = orc.g1.to_affine(ctx.test_op(5, op, p)[0])[0]
        assert [L.fp_to_int(aff["x"]), L.fp_to_int(aff["y"])] == v[2:4], op
    v = [H(x) for x in kats["g1_add"]["values"]]
    a = _aff_from_ints(L.G1_AFFINE, v[0], v[1]); b = _aff_from_ints(L.G1_AFFINE, v[2], v[3])
    for op in (0, 3):
        aff = orc.g1.to_affine(ctx.test_op(5, op, a, b)[0])[0]
        assert [L.fp_to_int(aff["x"]), L.fp_to_int(aff["y"])] == v[4:6], op


@pytest.mark.parametrize("family", [5, 6], ids=["g1", "g2"])
def test_group_law_vs_oracle_incl_exceptional_cases(ctx, orc, family):
    n = 40
    g = orc.g1 if family == 5 else orc.g2
    A = hg.g1_progression(71, 3, n) if family == 5 else hg.g2_progression(71, 3, n)
    B = hg.g1_progression(91, 7, n) if family == 5 else hg.g2_progression(91, 7, n)
    neg = hg.g1_neg if family == 5 else hg.g2_neg
    B[1] = A[1]                                                           # P + P (g1.go:430-437 falls back to Double)
    B[2] = neg(A[2:3])[0]                                                 # P + (-P) = O
    A["inf"][3] = 1                                                       # O + Q
    B["inf"][4] = 1                                                       # P + O
    A["inf"][5] = 1; B["inf"][5] = 1                                      # O + O
    JAC = L.G1_JAC if family == 5 else L.G2_JAC
    def jac_of(aff):
        j = np.zeros(aff.size, dtype=JAC)
        one = mont(1)
        for i in range(aff.size):
            if aff["inf"][i]:
                continue
            j["x"][i] = aff["x"][i]; j["y"][i] = aff["y"][i]
            if family == 5: j["z"][i] = one
            else: j["z"][i][0] = one
        return j
    want_add = g.to_affine(g.add(jac_of(A), jac_of(B)))
    want_dbl = g.to_affine(g.double(jac_of(A)))
    for op in (0, 3):
        got = g.to_affine(ctx.test_op(family, op, A, B)[0])
        assert got.tobytes() == want_add.tobytes(), op
    for op in (1, 4):
        got = g.to_affine(ctx.test_op(family, op, A)[0])
        assert got.tobytes() == want_dbl.tobytes(), op
    got = g.to_affine(ctx.test_op(family, 2, A, B)[0])
    assert got.tobytes() == g.to_affine(g.double(g.add(jac_of(A), jac_of(B)))).tobytes()
