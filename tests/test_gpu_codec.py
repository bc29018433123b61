"""GPU parity of the wire-format / scalar-multiplication entry points (include/b381.h: b381_g{1,2}_decompress_batch,
b381_g{1,2}_compress_batch, b381_g{1,2}_mul_batch) against the oracle's DecompressG1/G2, CompressG1/G2 and
MulFR + ToAffine (g1.go:59-90,185-249,322-340; g2.go:70-102,219-289,365-386), through the C ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import codec_cases as cc
from test_emu_codec import _expected_points

from bls_b200 import hostgen as hg, layout as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from bls_b200 import capi
    return capi.Ctx(0)


def _grp(orc, which):
    return (orc.g1, L.G1_AFFINE, 48) if which == "g1" else (orc.g2, L.G2_AFFINE, 96)


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_decompress_matches_oracle(ctx, orc, which):
    grp, dtype, nb = _grp(orc, which)
    pts = hg.g1_progression(5, 3, 40) if which == "g1" else hg.g2_progression(7, 5, 40)
    cases = cc.compressed_cases(orc, grp, nb, pts, 2) + [cc.REF_INVALID_G1 if which == "g1" else cc.REF_INVALID_G2]
    dec = ctx.g1_decompress_batch if which == "g1" else ctx.g2_decompress_batch
    for checked in (False, True):
        exp_p, exp_s = cc.oracle_decompress(grp, dtype, cases, checked)
        got, st = dec(b"".join(cases), check_subgroup=checked)
        assert st.tolist() == exp_s.tolist()
        assert got.tobytes() == _expected_points(exp_p, exp_s).tobytes()
    assert exp_s[-1] != 0 and set(exp_s.tolist()) >= {0, 1, 2, 3, 4}


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_compress_decompress_roundtrip_large(ctx, orc, which):
    """size-independent property at batch scale: decompress(compress(P)) == P with status 0 for subgroup points,
    and the compressed bytes equal the oracle's on a sample"""
    grp, dtype, nb = _grp(orc, which)
    n = 4096
    pts = hg.g1_progression(0xC0DEC, 0x9E3779B9, n) if which == "g1" else hg.g2_progression(0xC0DEC, 0x9E3779B9, n)
    comp = (ctx.g1_compress_batch if which == "g1" else ctx.g2_compress_batch)(pts)
    for i in range(0, n, 257):
        assert comp[i].tobytes() == grp.compress(pts[i:i + 1])
    back, st = (ctx.g1_decompress_batch if which == "g1" else ctx.g2_decompress_batch)(comp.tobytes(), check_subgroup=True)
    assert not st.any()
    assert back.tobytes() == pts.tobytes()


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_mul_batch(ctx, orc, which):
    grp, dtype, _ = _grp(orc, which)
    n = 96
    pts = hg.g1_progression(17, 3, n) if which == "g1" else hg.g2_progression(19, 5, n)
    k = orc.XorShift(78).rand_fr(n)
    k[0] = 0; k[1] = [1, 0, 0, 0]; k[2] = L.int_to_limbs(L.R_ORDER - 1, 4)
    mul = ctx.g1_mul_batch if which == "g1" else ctx.g2_mul_batch
    assert mul(pts, k).tobytes() == grp.to_affine(grp.mul_fr(pts, k, threads=8)).tobytes()
    # one base, many scalars (PrivToPub, g1pubs/bls.go:144-146): sk_i = s + i d gives the progression points
    s0, d0, m = 0x1234567, 0x89AB, 2048
    sk = np.array([L.int_to_limbs(s0 + i * d0, 4) for i in range(m)], np.uint64)
    gen = hg.g1_mul(1) if which == "g1" else hg.g2_mul(1)
    exp = hg.g1_progression(s0, d0, m) if which == "g1" else hg.g2_progression(s0, d0, m)
    assert mul(gen, sk).tobytes() == exp.tobytes()
    # many points, one scalar (Sign of many hashed messages with one key)
    got = mul(pts, k[5:6])
    assert got.tobytes() == grp.to_affine(grp.mul_fr(pts, np.repeat(k[5:6], n, axis=0), threads=8)).tobytes()
    # infinity in, infinity out
    z = np.zeros(1, dtype=dtype); z["inf"] = 1
    assert int(mul(z, k[3:4])["inf"][0]) == 1
    # the endomorphism ladders (b381_g{1,2}_mul_subgroup_batch): same bytes on points of the group, all stride forms
    smul = ctx.g1_mul_subgroup_batch if which == "g1" else ctx.g2_mul_subgroup_batch
    assert smul(pts, k).tobytes() == mul(pts, k).tobytes()
    assert smul(gen, sk).tobytes() == exp.tobytes()
    assert smul(pts, k[5:6]).tobytes() == got.tobytes()
    assert int(smul(z, k[3:4])["inf"][0]) == 1
    X = 0xd201000000010000
    edge = np.array([L.int_to_limbs(v, 4) for v in (X, X * X - 1, X ** 3, L.R_ORDER, X ** 4 + 7, (1 << 256) - 1)], np.uint64)
    assert smul(pts[:6], edge).tobytes() == mul(pts[:6], edge).tobytes()


def test_empty_batches(ctx):
    got, st = ctx.g1_decompress_batch(b"")
    assert got.size == 0 and st.size == 0
    assert ctx.g2_compress_batch(np.zeros(0, dtype=L.G2_AFFINE)).shape == (0, 96)
