"""CPU unit tests of the DEVICE logic: bls_b200/csrc/*.cuh compiled as plain C++ (tests/emu/emu.cc,
portable limb bodies) and compared with the oracle.  This exercises the exact tower / pairing /
group-law / MSM code the kernels run -- without a GPU -- so formula errors surface in the CPU suite.
The product library never contains this host build."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U64 = np.uint64


@pytest.fixture(scope="module")
def emu():
    import __graft_entry__ as g
    so = g.build_emu()
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_fp_ops(emu, orc):
    xs = orc.XorShift(11)
    a = xs.rand_fq(300); b = xs.rand_fq(300)
    q1 = np.array(L.int_to_limbs(L.Q - 1), U64)
    a[0] = 0; b[1] = 0; a[2] = q1; b[2] = q1; a[3] = q1; b[3] = L.fp_from_int(1)
    for op, name in [(0, "mul"), (1, "add"), (2, "sub"), (3, "square"), (4, "neg"), (5, "double")]:
        out = np.empty_like(a)
        emu.emu_fp_op(op, _p(a), _p(b), _p(out), ctypes.c_size_t(300))
        assert (out == orc.fq(name, a, b)).all(), name
    a[0] = L.fp_from_int(7)
    out = np.empty_like(a[:20])
    emu.emu_fp_op(6, _p(a), _p(b), _p(out), ctypes.c_size_t(20))     # fp_inv == the oracle's binary-Euclid inverse (fq.go:224-266)
    assert (out == orc.fq("inverse", a[:20])).all()


def test_fp_inv_almost_inverse(emu):
    """fp_inv (Kaliski's almost-inverse with batched shifts) on limb patterns that drive every branch -- 0, 1 (few shifts: the
    2^e > Q correction), powers of two (whole-limb shifts), Q - small, random -- against pow(-1) and the Fermat chain"""
    rng = np.random.RandomState(21)
    vals = [0, 1, 2, 3, L.Q - 1, L.Q - 2, (L.Q - 1) // 2, (L.Q + 1) // 2, 1 << 32, 1 << 64, 1 << 352, 1 << 380, (1 << 380) + 1,
            3 << 96, (1 << 381) - 1 - (1 << 200), 0xffffffff, 1 << 31] + [1 << k for k in range(5, 380, 17)]
    vals += [int.from_bytes(rng.bytes(48), "big") % L.Q for _ in range(400)]
    a = np.zeros((len(vals), 6), U64)
    for i, v in enumerate(vals):
        a[i] = np.array(L.int_to_limbs(v), dtype=U64)       # the limb integer A itself; the result is A^-1 R^2 mod Q
    for op in (6, 8):
        out = np.empty_like(a)
        emu.emu_fp_op(op, _p(a), _p(a), _p(out), ctypes.c_size_t(len(vals)))
        for i, v in enumerate(vals):
            want = 0 if v == 0 else pow(v, -1, L.Q) * L.MONT_R * L.MONT_R % L.Q
            assert L.limbs_to_int(out[i]) == want, (op, hex(v))


def test_fp12_ops(emu, orc):
    xs = orc.XorShift(12)
    n = 6
    a = xs.rand_fq(12 * n).reshape(n, 2, 3, 2, 6); b = xs.rand_fq(12 * n).reshape(n, 2, 3, 2, 6)
    out = np.empty_like(a)
    for op, name, arg in [(0, "mul", 0), (3, "square", 0), (6, "inverse", 0), (7, "frobenius", 1), (7, "frobenius", 2),
                          (7, "frobenius", 3), (12, "conjugate", 0)]:
        emu.emu_fp12_op(op, ctypes.c_uint64(arg), _p(a), _p(b), _p(out), ctypes.c_size_t(n))
        assert (out == orc.fq12(name, a, b, arg)).all(), (name, arg)
    # sparse multiplication: fq12.go:32-47 with (c0, c1, c4) taken from b
    emu.emu_fp12_op(13, ctypes.c_uint64(0), _p(a), _p(b), _p(out), ctypes.c_size_t(n))
    assert (out == orc.fq12("mul_by_014", a, b)).all()


def test_cyclotomic_square_and_exp_by_x(emu, orc):
    """valid only in the cyclotomic subgroup: use final-exponentiation outputs"""
    P = hg.g1_progression(3, 1, 2); Q = hg.g2_progression(4, 1, 2)
    f = orc.pairing_batch(P, Q)
    out = np.empty_like(f)
    emu.emu_fp12_op(15, ctypes.c_uint64(0), _p(f), _p(f), _p(out), ctypes.c_size_t(2))
    assert (out == orc.fq12("square", f)).all()
    emu.emu_fp12_op(16, ctypes.c_uint64(L.BLS_X), _p(f), _p(f), _p(out), ctypes.c_size_t(2))
    assert (out == orc.fq12("conjugate", orc.fq12("exp", f, None, L.BLS_X))).all()    # ExpByX, pairing.go:92-98
    # |x| / 2 (the second ExpByX of the chain) and an exponent outside the compressed schedule (Granger-Scott fallback)
    for x in (L.BLS_X >> 1, 0x1234567):
        emu.emu_fp12_op(16, ctypes.c_uint64(x), _p(f), _p(f), _p(out), ctypes.c_size_t(2))
        assert (out == orc.fq12("conjugate", orc.fq12("exp", f, None, x))).all()
    # the degenerate value 1 (every compressed coordinate is zero; what a pairing with a point at infinity produces): the
    # decompression would divide by zero, the fallback must give 1
    one = np.zeros_like(f[:1]); one[0, 0, 0, 0] = L.fp_from_int(1)
    for v in (one,):
        o1 = np.empty_like(v)
        emu.emu_fp12_op(16, ctypes.c_uint64(L.BLS_X), _p(v), _p(v), _p(o1), ctypes.c_size_t(1))
        assert (o1 == orc.fq12("conjugate", orc.fq12("exp", v, None, L.BLS_X))).all()


def test_miller_loop_and_final_exp(emu, orc, kats):
    P = np.concatenate([orc.g1_generator(), hg.g1_progression(0x99, 7, 3)])
    Q = np.concatenate([orc.g2_generator(), hg.g2_progression(0x55, 9, 3)])
    ml = np.zeros(4, dtype=L.FP12)
    emu.emu_miller_loop(_p(P), _p(Q), ctypes.c_size_t(4), _p(ml))
    for i in range(4):
        assert (ml[i] == orc.miller_loop(P[i:i + 1], Q[i:i + 1])).all()
    fe = np.zeros(4, dtype=L.FP12); ok = np.zeros(4, np.uint8)
    emu.emu_final_exp(_p(ml), ctypes.c_size_t(4), _p(fe), _p(ok))
    assert ok.all() and fe.tobytes() == orc.pairing_batch(P, Q).tobytes()
    exp = np.stack([L.fp_from_int(int(x, 16)) for x in kats["pairing_g1_g2"]["coeffs"]]).reshape(2, 3, 2, 6)
    assert (fe[0] == exp).all()                                   # RELIC vector, pairing_test.go:9-58


def test_miller_loop_two_pairs_shared_accumulator(emu, orc):
    """miller_loop_two (pairing.go:16-75 with two items, the core of CompareTwoPairings pairing.go:140-147): equal to the
    oracle's multi-item MillerLoop when it exposes one, and always to the product of the two single-pair Miller values;
    an infinity on either side of a pair drops that pair"""
    P = hg.g1_progression(0x31, 5, 6); Q = hg.g2_progression(0x47, 3, 6)
    P["inf"][4] = 1                                                 # groups: (0,1) (2,3) (4,5) with pair 4 at infinity
    out = np.zeros(3, dtype=L.FP12)
    emu.emu_miller_loop2(_p(P), _p(Q), ctypes.c_size_t(3), _p(out))
    for g in range(3):
        a = orc.miller_loop(P[2 * g:2 * g + 1], Q[2 * g:2 * g + 1]) if not P["inf"][2 * g] else None
        b = orc.miller_loop(P[2 * g + 1:2 * g + 2], Q[2 * g + 1:2 * g + 2])
        exp = b if a is None else orc.fq12("mul", a, b)
        assert (out[g] == exp).all(), g


def test_group_sums(emu, orc):
    P = hg.g1_progression(21, 5, 40)
    P = np.concatenate([P, P[:3], hg.g1_neg(P[5:7])]); P["inf"][9] = 1
    out = np.zeros(1, dtype=L.G1_JAC)
    for fn in ("emu_g1_sum", "emu_g1_sum_tree"):
        getattr(emu, fn)(_p(P), ctypes.c_size_t(P.size), _p(out))
        assert orc.g1.to_affine(out).tobytes() == orc.g1.to_affine(orc.g1.sum_affine(P)).tobytes(), fn
    Q = hg.g2_progression(22, 3, 17); Q = np.concatenate([Q, Q[:2]])
    o2 = np.zeros(1, dtype=L.G2_JAC)
    emu.emu_g2_sum(_p(Q), ctypes.c_size_t(Q.size), _p(o2))
    assert orc.g2.to_affine(o2).tobytes() == orc.g2.to_affine(orc.g2.sum_affine(Q)).tobytes()


@pytest.mark.parametrize("rank,nranks", [(0, 1), (0, 2), (1, 2)])
def test_msm_building_blocks(emu, orc, rank, nranks):
    n = 200
    P = hg.g1_progression(31, 2, n)
    K, _ = hg.splitmix_scalars(5, n)
    out = np.zeros(1, dtype=L.G1_JAC)
    emu.emu_g1_msm(_p(P), _p(K), ctypes.c_size_t(n), 6, rank, nranks, 16, _p(out))
    if nranks == 1:
        assert orc.g1.to_affine(out).tobytes() == orc.g1.to_affine(orc.g1_msm_naive(P, K, threads=4)).tobytes()
    else:
        other = np.zeros(1, dtype=L.G1_JAC)
        emu.emu_g1_msm(_p(P), _p(K), ctypes.c_size_t(n), 6, 1 - rank, nranks, 16, _p(other))
        tot = np.zeros(1, dtype=L.G1_JAC)
        emu.emu_g1_fold(_p(np.concatenate([out, other])), ctypes.c_size_t(2), _p(tot))
        assert orc.g1.to_affine(tot).tobytes() == orc.g1.to_affine(orc.g1_msm_naive(P, K, threads=4)).tobytes()


def test_op_counts(emu):
    """The multiplier work bench.py's roofline numerator uses, pinned by the instrumented host build of the device code: Fq
    products (300 wide MACs: 144 product + 144 reduction + 12 quotient words) and two-product dot products with one reduction
    (444 wide MACs) executed per pairing by the one-pairing-per-thread kernels.  The two-lane form executes the same
    operations (the host emulation runs two copies of the lane pair); the four-lane form trades a few more for balance."""
    import bench
    emu.emu_mul_count.restype = ctypes.c_ulonglong; emu.emu_dot2_count.restype = ctypes.c_ulonglong
    P = hg.g1_progression(3, 1, 1); Q = hg.g2_progression(4, 1, 1)
    ml = np.zeros(1, dtype=L.FP12); fe = np.zeros(1, dtype=L.FP12); ok = np.zeros(1, np.uint8)

    def macs():
        m = emu.emu_mul_count(1); d = emu.emu_dot2_count(1)
        return 300 * (m - 2 * d) + 444 * d
    macs()
    emu.emu_miller_loop(_p(P), _p(Q), ctypes.c_size_t(1), _p(ml)); a = macs()
    emu.emu_final_exp(_p(ml), ctypes.c_size_t(1), _p(fe), _p(ok)); b = macs()
    assert (a, b) == (bench.MACS_IMPL_MILLER, bench.MACS_IMPL_FINAL_EXP) == (2052576, 1895976)
    emu.emu_duo_miller_loop(_p(P), _p(Q), ctypes.c_size_t(1), _p(ml)); a2 = macs()
    emu.emu_duo_final_exp(_p(ml), ctypes.c_size_t(1), _p(fe), _p(ok)); b2 = macs()
    # two-lane form: the Fq2 squarings go through the two-product body (444 instead of 300 wide MACs per lane; keeps the plain
    # multiplier out of the hot instruction set), and both lanes of a pair run the six norm inversions (+ 12 products)
    assert (a2, b2) == (2 * 2209248, 2 * 1993752)
    P2 = hg.g1_progression(3, 1, 2); Q2 = hg.g2_progression(4, 1, 2)
    emu.emu_miller_loop2(_p(P2), _p(Q2), ctypes.c_size_t(1), _p(ml)); c = macs()
    assert c == bench.MACS_IMPL_MILLER2


def test_prepared_g2(emu, orc):
    """g2_prepare_one == the oracle's G2AffineToPrepared coefficients (g2.go:650-801: 68 triples in the reference's order),
    miller_loop_prepared_one == MillerLoop on the unprepared point, the fused + prepared two-pair loop == miller_loop_two"""
    n = 4
    P = hg.g1_progression(0x71, 3, 2 * n); Q = hg.g2_progression(0x33, 5, n)
    Q["inf"][3] = 1
    prep = np.zeros(n, dtype=L.G2_PREPARED)
    emu.emu_g2_prepare(_p(Q), ctypes.c_size_t(n), _p(prep))
    for i in range(3):
        want = orc.g2_prepare(Q[i:i + 1])
        assert want.shape[0] == 68 and (prep["coeffs"][i] == want).all(), i
    assert prep["inf"].tolist() == [0, 0, 0, 1] and not prep["coeffs"][3].any()
    idx = np.array([0, 1, 2, 3, 0, 0, 2, 1], np.uint32)
    out = np.zeros(2 * n, dtype=L.FP12)
    emu.emu_miller_loop_prepared(_p(P), _p(prep), _p(idx), ctypes.c_size_t(2 * n), _p(out))
    ref = np.zeros(2 * n, dtype=L.FP12)
    Qi = Q[idx]
    emu.emu_miller_loop(_p(P), _p(Qi), ctypes.c_size_t(2 * n), _p(ref))
    assert out.tobytes() == ref.tobytes()
    for i in (0, 1, 2):
        assert (out[i] == orc.miller_loop(P[i:i + 1], Qi[i:i + 1])).all()
    # groups of two pairs: (P[2g], Q0[g]) computed, (P[2g+1], prep[gidx[g]]) from coefficients
    gidx = np.array([1, 0, 3, 2], np.uint32)
    Q0 = hg.g2_progression(0x99, 7, n); Q0["inf"][1] = 1
    QQ = np.zeros(2 * n, dtype=L.G2_AFFINE); QQ[0::2] = Q0; QQ[1::2] = Q[gidx]
    got = np.zeros(n, dtype=L.FP12); want2 = np.zeros(n, dtype=L.FP12)
    emu.emu_miller_loop_fused_prepared(_p(P), _p(Q0), _p(prep), _p(gidx), ctypes.c_size_t(n), _p(got))
    emu.emu_miller_loop2(_p(P), _p(QQ), ctypes.c_size_t(n), _p(want2))
    assert got.tobytes() == want2.tobytes()
