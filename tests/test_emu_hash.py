"""CPU tests of csrc/hash.cuh (host build in tests/emu): SHA-256 against hashlib and HashG2WithDomain against the
reference's known answer (hash_test.go:72-82) and the host restatement pinned by it (oracle/hostmath.py)."""
import ctypes
import hashlib
import os
import sys

import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L
from oracle import hostmath as hm


@pytest.fixture(scope="module")
def emu():
    import __graft_entry__ as g
    return ctypes.CDLL(g.build_emu())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hash_cases(n, seed):
    rng = np.random.RandomState(seed)
    msgs = [bytes(32), b"\xff" * 32] + [rng.randint(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(n - 2)]
    doms = [bytes(8), b"\xff" * 8] + [rng.randint(0, 256, 8, dtype=np.uint8).tobytes() for _ in range(n - 2)]
    return msgs, doms


def expected_hashes(msgs, doms):
    return hg.g2_points([hm.hash_g2_with_domain(m, d) for m, d in zip(msgs, doms)])


def test_sha256_short(emu):
    rng = np.random.RandomState(3)
    for ln in (0, 1, 31, 41, 55):
        m = rng.randint(0, 256, max(ln, 1), dtype=np.uint8)
        dg = np.zeros(8, np.uint32)
        emu.emu_sha256_short(_p(m), ln, _p(dg))
        assert b"".join(int(w).to_bytes(4, "big") for w in dg) == hashlib.sha256(m[:ln].tobytes()).digest()


def test_hash_g2_with_domain_kat_and_host(emu, orc, kats):
    msgs, doms = hash_cases(6, 4)
    m = np.frombuffer(b"".join(msgs), np.uint8).copy(); d = np.frombuffer(b"".join(doms), np.uint8).copy()
    out = np.zeros(len(msgs), dtype=L.G2_AFFINE)
    emu.emu_hash_g2_with_domain(_p(m), _p(d), ctypes.c_size_t(1), ctypes.c_size_t(len(msgs)), _p(out))
    # hash_test.go:72-82: CompressG2(HashG2WithDomain(0^32, 0^8).ToAffine())
    assert orc.g2.compress(out[:1]).hex() == kats["hash"]["hash_g2_with_domain_zero_compressed"]
    assert out.tobytes() == expected_hashes(msgs, doms).tobytes()
    for i in range(len(msgs)):
        assert orc.g2.in_subgroup(out[i:i + 1])
    # one domain for the whole batch
    out1 = np.zeros(len(msgs), dtype=L.G2_AFFINE)
    emu.emu_hash_g2_with_domain(_p(m), _p(d), ctypes.c_size_t(0), ctypes.c_size_t(len(msgs)), _p(out1))
    assert out1.tobytes() == expected_hashes(msgs, [doms[0]] * len(msgs)).tobytes()


def swu_messages():
    rng = np.random.RandomState(8)
    return [b"the message to be signed", b"", b"a", bytes(range(54)), bytes(range(55)), bytes(range(56)), bytes(range(64)),
            rng.bytes(119), rng.bytes(200)]


def pack(msgs):
    off = np.zeros(len(msgs) + 1, np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    raw = np.frombuffer(b"".join(msgs) + b"\0", np.uint8).copy()
    return raw, off


def test_sha256_prefixed(emu):
    for m in swu_messages():
        raw = np.frombuffer(m + b"\0", np.uint8).copy()
        dg = np.zeros(8, np.uint32)
        emu.emu_sha256_prefixed(ctypes.c_uint8(1), _p(raw), ctypes.c_size_t(len(m)), _p(dg))
        assert b"".join(int(w).to_bytes(4, "big") for w in dg) == hashlib.sha256(b"\x01" + m).digest()


def test_hash_g1_g2_kats_and_host(emu, orc, kats):
    """HashG1 / HashG2 (hash.go:320-331,404-411): the reference's known answers (hash_test.go:12-26,48-62) and the
    pinned host restatement on messages of every padding class"""
    msgs = swu_messages()
    raw, off = pack(msgs)
    o1 = np.zeros(len(msgs), dtype=L.G1_AFFINE); o2 = np.zeros(len(msgs), dtype=L.G2_AFFINE)
    emu.emu_hash_g1(_p(raw), _p(off), ctypes.c_size_t(len(msgs)), _p(o1))
    emu.emu_hash_g2(_p(raw), _p(off), ctypes.c_size_t(len(msgs)), _p(o2))
    h = kats["hash"]
    assert msgs[0].decode() == h["message"]
    assert [hex(L.fp_to_int(o1["x"][0])), hex(L.fp_to_int(o1["y"][0]))] == h["hash_g1"]
    assert o1.tobytes() == hg.g1_points([hm.hash_g1(m) for m in msgs]).tobytes()
    assert o2.tobytes() == hg.g2_points([hm.hash_g2(m) for m in msgs]).tobytes()
    for i in range(len(msgs)):
        assert orc.g1.in_subgroup(o1[i:i + 1]) and orc.g2.in_subgroup(o2[i:i + 1])


def test_scale_by_cofactor_endomorphism_ladder(emu):
    """[h2] P (ScaleByCofactor, g2.go:133,1041-1085) through clearH2 and the base-|x| psi ladder equals the plain 507-bit
    ladder and the host restatement on curve points outside G2, inside G2 and inside the cofactor subgroup"""
    rng = np.random.RandomState(77)
    pts = []
    while len(pts) < 4:
        x = (int.from_bytes(rng.bytes(47), "big"), int.from_bytes(rng.bytes(47), "big"))
        y = hm.fq2_sqrt(hm._Fq2.add(hm._Fq2.mul(hm._Fq2.sqr(x), x), (4, 4)))
        if y is not None:
            pts.append((x, y))
    torsion = hm.g2_mul(pts[0], hm.G2_COFACTOR)                  # a point of G2
    low = hm.g2_mul(pts[1], hm.R_ORDER)                           # order divides h2: [h2] low = O
    assert hm.g2_in_subgroup(torsion) and low is not None
    pts += [torsion, hm.g2_neg(torsion), low]
    src = hg.g2_points(pts)
    want = hg.g2_points([hm.g2_mul(p, hm.G2_COFACTOR) for p in pts])
    assert want["inf"][-1] == 1
    for which in (0, 1):
        out = np.zeros(len(pts), dtype=L.G2_AFFINE)
        emu.emu_g2_scale_by_cofactor(which, _p(src), ctypes.c_size_t(len(pts)), _p(out))
        assert out.tobytes() == want.tobytes(), which


def test_fp2_sqrt_norm_method(emu):
    """fp2_sqrt (two Fq exponentiations through the norm) against FQ2.Sqrt (fq2.go:198-232) as the host restatement and
    the device's own Algorithm 9: same verdict, and the same root up to the sign every caller normalises"""
    rng = np.random.RandomState(5)
    rnd = lambda: int.from_bytes(rng.bytes(47), "big")
    vals = [(rnd(), rnd()) for _ in range(12)]
    vals += [hm._Fq2.sqr(v) for v in vals[:4]]                                # certain squares
    vals += [(0, 0), (1, 0), (hm.Q - 1, 0), (4, 0), (5, 0), (0, 1), (0, hm.Q - 1), (0, rnd()), (rnd(), 0), (3, 4)]
    a = np.zeros((len(vals), 12), np.uint64)
    for i, (c0, c1) in enumerate(vals):
        a[i, :6] = L.fp_from_int(c0); a[i, 6:] = L.fp_from_int(c1)
    res = []
    for which in (0, 1):
        out = np.zeros_like(a); ok = np.zeros(len(vals), np.uint8)
        emu.emu_fp2_sqrt(which, _p(a), ctypes.c_size_t(len(vals)), _p(out), _p(ok))
        res.append((out, ok))
    assert res[0][1].tolist() == res[1][1].tolist()
    seen = set()
    for i, v in enumerate(vals):
        want = hm.fq2_sqrt(v)
        assert (want is not None) == bool(res[0][1][i]), v
        if want is None:
            continue
        got = (L.fp_to_int(res[0][0][i, :6]), L.fp_to_int(res[0][0][i, 6:]))
        ref = (L.fp_to_int(res[1][0][i, :6]), L.fp_to_int(res[1][0][i, 6:]))
        assert ref == want
        assert got in (want, hm.fq2_neg(want)), v
        assert hm._Fq2.sqr(got) == (v[0] % hm.Q, v[1] % hm.Q)
        seen.add(got == want)
    assert len(vals) - int(res[0][1].sum()) >= 3                             # non-squares were exercised


def test_fp_is_square_jacobi(emu):
    """the binary Jacobi symbol against Euler's criterion on random values, small values, powers of two and Q - small"""
    rng = np.random.RandomState(9)
    vals = [int.from_bytes(rng.bytes(48), "big") % hm.Q for _ in range(300)]
    vals += list(range(0, 40)) + [hm.Q - k for k in range(1, 40)] + [1 << k for k in range(0, 381, 7)]
    vals += [(1 << 64) * 3, (1 << 352) + (1 << 32), (hm.Q - 1) // 2, (hm.Q + 1) // 2]
    a = np.zeros((len(vals), 6), np.uint64)
    for i, v in enumerate(vals):
        a[i] = np.array(L.int_to_limbs(v), dtype=np.uint64)      # the symbol of the limb integer itself (chi(2^384) = 1)
    out = np.zeros(len(vals), np.uint8)
    emu.emu_fp_is_square(_p(a), ctypes.c_size_t(len(vals)), _p(out))
    want = [1 if v % hm.Q == 0 or pow(v, (hm.Q - 1) // 2, hm.Q) == 1 else 0 for v in vals]
    assert out.tolist() == want
    assert 100 < sum(want) < len(vals) - 100
