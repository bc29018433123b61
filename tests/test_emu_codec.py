"""CPU tests of csrc/codec.cuh (compiled as host C++ in tests/emu) against the oracle: DecompressG1/G2 (checked and
unchecked, g1.go:185-227, g2.go:219-265), CompressG1/G2 (g1.go:230-249, g2.go:268-289), MulFR + ToAffine."""
import ctypes

import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import codec_cases as cc


@pytest.fixture(scope="module")
def emu():
    import __graft_entry__ as g
    return ctypes.CDLL(g.build_emu())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _expected_points(pts, st):
    """the engine writes the canonical zero for statuses 1-3; the oracle leaves its output untouched there"""
    out = pts.copy()
    for i in np.nonzero((st != 0) & (st != 4))[0]:
        out[i] = np.zeros(1, dtype=pts.dtype)[0]
        out["y"][i] = L.fp_from_int(1) if pts.dtype == L.G1_AFFINE else np.array([L.fp_from_int(1), L.fp_from_int(0)])
        out["inf"][i] = 1
    return out


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_decompress_matches_oracle(emu, orc, which):
    grp, dtype, nb = (orc.g1, L.G1_AFFINE, 48) if which == "g1" else (orc.g2, L.G2_AFFINE, 96)
    pts = hg.g1_progression(5, 3, 6) if which == "g1" else hg.g2_progression(7, 5, 6)
    cases = cc.compressed_cases(orc, grp, nb, pts, 1) + [cc.REF_INVALID_G1 if which == "g1" else cc.REF_INVALID_G2]
    raw = np.frombuffer(b"".join(cases), np.uint8).copy()
    n = len(cases)
    for checked in (0, 1):
        exp_p, exp_s = cc.oracle_decompress(grp, dtype, cases, bool(checked))
        got = np.zeros(n, dtype=dtype); st = np.zeros(n, np.uint8)
        getattr(emu, "emu_%s_decompress" % which)(_p(raw), ctypes.c_size_t(n), checked, _p(got), _p(st))
        assert st.tolist() == exp_s.tolist()
        assert got.tobytes() == _expected_points(exp_p, exp_s).tobytes()
    assert exp_s[-1] != 0, "the reference's invalid public key must be rejected (bls_test.go TestPubkeyDeserializeInvalid)"
    assert set(exp_s.tolist()) >= {0, 1, 2, 3, 4}, "every error path is exercised"


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_compress_roundtrip(emu, orc, which):
    grp, dtype, nb = (orc.g1, L.G1_AFFINE, 48) if which == "g1" else (orc.g2, L.G2_AFFINE, 96)
    pts = hg.g1_progression(11, 7, 9) if which == "g1" else hg.g2_progression(13, 9, 9)
    pts = np.concatenate([pts, hg.g1_neg(pts[:3]) if which == "g1" else pts[:0]])
    zero = np.zeros(1, dtype=dtype); zero["inf"] = 1; zero["y"] = L.fp_from_int(1) if which == "g1" else np.array([L.fp_from_int(1), L.fp_from_int(0)])
    pts = np.concatenate([pts, zero])
    out = np.zeros((pts.size, nb), np.uint8)
    getattr(emu, "emu_%s_compress" % which)(_p(pts), ctypes.c_size_t(pts.size), _p(out))
    for i in range(pts.size):
        assert out[i].tobytes() == grp.compress(pts[i:i + 1])


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_scalar_mul(emu, orc, which):
    grp, dtype = (orc.g1, L.G1_AFFINE) if which == "g1" else (orc.g2, L.G2_AFFINE)
    n = 6
    pts = hg.g1_progression(17, 3, n) if which == "g1" else hg.g2_progression(19, 5, n)
    xs = orc.XorShift(77)
    k = xs.rand_fr(n)
    k[0] = 0; k[1] = [1, 0, 0, 0]; k[2] = L.int_to_limbs(L.R_ORDER - 1, 4)
    exp = grp.to_affine(grp.mul_fr(pts, k))
    got = np.zeros(n, dtype=dtype)
    getattr(emu, "emu_%s_mul" % which)(_p(pts), ctypes.c_size_t(1), _p(k), ctypes.c_size_t(1), ctypes.c_size_t(n), _p(got))
    assert got.tobytes() == exp.tobytes()
    # one base, many scalars (PrivToPub) and one scalar, many points (Sign)
    got1 = np.zeros(n, dtype=dtype)
    getattr(emu, "emu_%s_mul" % which)(_p(pts), ctypes.c_size_t(0), _p(k), ctypes.c_size_t(1), ctypes.c_size_t(n), _p(got1))
    assert got1.tobytes() == grp.to_affine(grp.mul_fr(np.repeat(pts[:1], n), k)).tobytes()
    getattr(emu, "emu_%s_mul" % which)(_p(pts), ctypes.c_size_t(1), _p(k[3:]), ctypes.c_size_t(0), ctypes.c_size_t(n), _p(got1))
    assert got1.tobytes() == grp.to_affine(grp.mul_fr(pts, np.repeat(k[3:4], n, axis=0))).tobytes()


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_scalar_mul_endomorphism_ladder(emu, orc, which):
    """the GLV / psi ladders of b381_g{1,2}_mul_subgroup_batch on points of the group: same affine result as MulFR + ToAffine
    for scalars that exercise every digit pattern (0, 1, r - 1, |x|^i, |x|^i - 1, digits all ones, values >= r and >= |x|^4)"""
    grp, dtype = (orc.g1, L.G1_AFFINE) if which == "g1" else (orc.g2, L.G2_AFFINE)
    X = 0xd201000000010000
    ks = [0, 1, 2, L.R_ORDER - 1, L.R_ORDER - 2, X, X - 1, X + 1, X * X, X * X - 1, X ** 3, X ** 3 - 1, X ** 3 + X * X + X + 1,
          (X - 1) * (1 + X + X * X + X ** 3), L.R_ORDER, L.R_ORDER + 5, X ** 4, X ** 4 + 7, (1 << 256) - 1, (1 << 255) + 12345,
          0xffffffffffffffff, 1 << 64, (1 << 128) - 1, 1 << 128]
    rnd = orc.XorShift(91).rand_fr(12)
    k = np.concatenate([np.array([L.int_to_limbs(v, 4) for v in ks], np.uint64), rnd])
    n = k.shape[0]
    pts = hg.g1_progression(23, 7, n) if which == "g1" else hg.g2_progression(29, 11, n)
    plain = np.zeros(n, dtype=dtype); endo = np.zeros(n, dtype=dtype)
    getattr(emu, "emu_%s_mul" % which)(_p(pts), ctypes.c_size_t(1), _p(k), ctypes.c_size_t(1), ctypes.c_size_t(n), _p(plain))
    getattr(emu, "emu_%s_mul_subgroup" % which)(_p(pts), ctypes.c_size_t(1), _p(k), ctypes.c_size_t(1), ctypes.c_size_t(n), _p(endo))
    assert endo.tobytes() == plain.tobytes()
    canon = [i for i in range(n) if L.limbs_to_int(k[i]) < L.R_ORDER]
    assert endo[canon].tobytes() == grp.to_affine(grp.mul_fr(pts[canon], k[canon])).tobytes()
    assert int(endo["inf"][0]) == 1 and int(endo["inf"][ks.index(L.R_ORDER)]) == 1
    z = np.zeros(1, dtype=dtype); z["inf"] = 1; o = np.zeros(1, dtype=dtype)
    getattr(emu, "emu_%s_mul_subgroup" % which)(_p(z), ctypes.c_size_t(1), _p(k[5:]), ctypes.c_size_t(1), ctypes.c_size_t(1), _p(o))
    assert int(o["inf"][0]) == 1


def test_scalar_mul_small_order_points(emu, orc):
    """the windowed ladder of b381_g{1,2}_mul_batch on points whose multiples hit infinity inside the table (order 3 on E(Fq),
    13 on E'(Fq2)): same result as the reference's double-and-add"""
    from oracle import hostmath as hm
    rng = np.random.RandomState(3)
    h1 = 0x396c8c005555e1568c00aaab0000aaab
    g1pts, g2pts = [], []
    while len(g1pts) < 2:
        x = int.from_bytes(rng.bytes(47), "big")
        y = hm.fq_sqrt((x * x * x + 4) % L.Q)
        if y is None:
            continue
        p = hm.g1_mul((x, y), h1 * L.R_ORDER // 3)
        if p is not None and hm.g1_mul(p, 3) is None:
            g1pts.append(p)
    while len(g2pts) < 1:
        x = (int.from_bytes(rng.bytes(47), "big"), int.from_bytes(rng.bytes(47), "big"))
        y = hm.fq2_sqrt(hm._Fq2.add(hm._Fq2.mul(hm._Fq2.sqr(x), x), (4, 4)))
        if y is None:
            continue
        p = hm.g2_mul((x, y), hm.G2_COFACTOR * L.R_ORDER // 169)       # the 13-part of E'(Fq2) is Z_13 x Z_13
        if p is not None and hm.g2_mul(p, 13) is None:
            g2pts.append(p)
    ks = [1, 2, 3, 10, 11, 12, 13, 26, 0x123456789abcdef, L.R_ORDER - 1]
    for which, pts, mul, conv in (("g1", g1pts, hm.g1_mul, hg.g1_points), ("g2", g2pts, hm.g2_mul, hg.g2_points)):
        src = conv([p for p in pts for _ in ks])
        k = np.array([L.int_to_limbs(v, 4) for _ in pts for v in ks], np.uint64)
        got = np.zeros(len(src), dtype=src.dtype)
        getattr(emu, "emu_%s_mul" % which)(_p(src), ctypes.c_size_t(1), _p(k), ctypes.c_size_t(1), ctypes.c_size_t(len(src)), _p(got))
        want = conv([mul(p, v) for p in pts for v in ks])
        assert got.tobytes() == want.tobytes(), which
        assert want["inf"].sum() >= 2


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_subgroup_criterion_on_cofactor_points(emu, orc, which):
    """The endomorphism membership tests (codec.cuh) must agree with the reference's [r]P == O (g1.go:137-141,
    g2.go:293-295) on every curve point -- in particular on points of the cofactor subgroups, of small order, and on
    sums of a subgroup point and a cofactor point."""
    from oracle import hostmath as hm
    grp, dtype, nb = (orc.g1, L.G1_AFFINE, 48) if which == "g1" else (orc.g2, L.G2_AFFINE, 96)
    mul, add = (hm.g1_mul, hm.g1_add) if which == "g1" else (hm.g2_mul, hm.g2_add)
    to_pods = hg.g1_points if which == "g1" else hg.g2_points
    gen = hm.G1 if which == "g1" else hm.G2
    cof = 76329603384216526031706109802092473003 if which == "g1" else hm.G2_COFACTOR   # g1.go:144, g2.go:133
    # random curve points: decompress random x values without the subgroup check
    rng = np.random.RandomState(11)
    pts = []
    while len(pts) < 4:
        b = bytearray(rng.randint(0, 256, nb, dtype=np.uint8).tobytes())
        b[0] = (b[0] & 0x1f) | 0x80
        if nb == 96:
            b[48] &= 0x1f
        e, o = grp.decompress(bytes(b), False)
        if e == 0:
            x = L.fp_to_int(o["x"][0]) if which == "g1" else (L.fp_to_int(o["x"][0][0]), L.fp_to_int(o["x"][0][1]))
            y = L.fp_to_int(o["y"][0]) if which == "g1" else (L.fp_to_int(o["y"][0][0]), L.fp_to_int(o["y"][0][1]))
            pts.append((x, y))
    small = [3, 11, 10177] if which == "g1" else [13, 23, 2713]
    cands = []
    for R in pts:
        T = mul(R, L.R_ORDER)                          # in the cofactor subgroup
        S = mul(R, cof)                                # in the r-torsion
        cands += [R, T, S, add(T, S), add(S, mul(gen, 5))]
        for p in small:
            if cof % p == 0:
                U = mul(T, cof // p)                   # order 1 or p
                if U is not None:
                    cands += [U, add(U, S)]
    cands = [c for c in cands if c is not None]
    arr = to_pods(cands)
    comp = b"".join(grp.compress(arr[i:i + 1]) for i in range(arr.size))
    raw = np.frombuffer(comp, np.uint8).copy()
    got = np.zeros(arr.size, dtype=dtype); st = np.zeros(arr.size, np.uint8)
    getattr(emu, "emu_%s_decompress" % which)(_p(raw), ctypes.c_size_t(arr.size), 1, _p(got), _p(st))
    exp = [0 if grp.in_subgroup(arr[i:i + 1]) else 4 for i in range(arr.size)]
    assert st.tolist() == exp
    assert 0 in exp and 4 in exp
