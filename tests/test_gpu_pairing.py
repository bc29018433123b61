"""Parity of the CUDA pairing path against the CPU oracle, through the C ABI (libb381.so).
Bit-exact: every Fq12 coefficient limb must match (integer arithmetic, no tolerance)."""
import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L

pytestmark = pytest.mark.gpu


# every test of this module runs once per schedule of the pairing arithmetic (b381_set_kernel_path): "auto" picks by batch size
# (four lanes per pairing at these sizes), the others force the warp-cooperative VM, the one-pairing-per-thread kernels of the
# 2^16 benchmark (k_miller_loop / k_final_exp / k_group_product), the two-lane kernels (k_duo_*), the four-lane kernels (k_quad_*)
@pytest.fixture(scope="module", params=["auto", "vm", "thread", "duo", "quad"])
def ctx(request):
    from bls_b200 import capi
    c = capi.Ctx(0, path=request.param)
    yield c
    c.close()


def _mont(v):
    return L.fp_from_int(v)


def test_pairing_kat_relic(ctx, kats, orc):
    """bls.Pairing(G1One, G2One) against the RELIC vector of pairing_test.go:9-58"""
    exp = np.stack([_mont(int(x, 16)) for x in kats["pairing_g1_g2"]["coeffs"]]).reshape(2, 3, 2, 6)
    out = ctx.pairing_batch(orc.g1_generator(), orc.g2_generator())
    assert (out[0] == exp).all()


@pytest.mark.parametrize("n", [1, 2, 31, 129, 300])
def test_pairing_batch_vs_oracle(ctx, orc, n):
    """ragged batch sizes (not multiples of the block / warp size)"""
    P = hg.g1_progression(0xB2000002 + n, 0x1234567, n)
    Q = hg.g2_progression(0x5EED + n, 0x7654321, n)
    got = ctx.pairing_batch(P, Q)
    exp = orc.pairing_batch(P, Q, threads=8)
    assert got.tobytes() == exp.tobytes()


def test_miller_loop_and_final_exp_checkpoints(ctx, orc):
    """pre-final-exponentiation value is bit-identical too (same step formulas as g2.go:655-772)"""
    n = 40
    P = hg.g1_progression(77, 5, n); Q = hg.g2_progression(99, 3, n)
    ml = ctx.miller_loop_batch(P, Q)
    for i in range(n):
        assert (ml[i] == orc.miller_loop(P[i:i + 1], Q[i:i + 1])).all(), i
    fe, ok = ctx.final_exp_batch(ml)
    assert ok.all()
    assert fe.tobytes() == orc.pairing_batch(P, Q, threads=8).tobytes()


def test_final_exp_of_zero_and_random(ctx, orc):
    """FinalExponentiation(0) is nil in the reference (pairing.go:83-85) -> ok = 0; random Fq12 inputs
    (not Miller outputs) exercise the easy part + cyclotomic squarings on arbitrary elements"""
    xs = orc.XorShift(3)
    f = xs.rand_fq(12 * 9).reshape(9, 2, 3, 2, 6)
    f[4] = 0
    fe, ok = ctx.final_exp_batch(f)
    assert ok.tolist() == [1, 1, 1, 1, 0, 1, 1, 1, 1]
    for i in range(9):
        good, exp = orc.final_exp(f[i])
        assert good == bool(ok[i])
        if good:
            assert (fe[i] == exp).all(), i


def test_pairing_batch_stream_matches_single_batches(ctx, orc):
    """b381_pairing_batch_stream (double-buffered host streaming) == b381_pairing_batch == the oracle, for batch sizes that
    divide n, leave a short last batch, exceed n, and for a single pair per batch"""
    n = 45
    P = hg.g1_progression(0x51, 7, n); Q = hg.g2_progression(0x29, 3, n)
    P["inf"][4] = 1
    want = ctx.pairing_batch(P, Q)
    assert want.tobytes() == orc.pairing_batch(P, Q).tobytes()
    for batch in (9, 16, 44, 45, 100, 1):
        assert ctx.pairing_batch_stream(P, Q, batch).tobytes() == want.tobytes(), batch
    assert ctx.pairing_batch_stream(P[:0], Q[:0], 8).size == 0
    with pytest.raises(Exception):
        ctx.pairing_batch_stream(P, Q, 0)


def test_empty_batch(ctx):
    P = np.zeros(0, dtype=L.G1_AFFINE); Q = np.zeros(0, dtype=L.G2_AFFINE)
    assert ctx.pairing_batch(P, Q).shape[0] == 0
    assert ctx.pairing_product_is_one(P, Q, [0]).shape[0] == 0


def test_infinity_pairs_contribute_one(ctx, orc):
    """documented divergence from SURVEY Q1: the reference panics, the engine returns the factor 1"""
    P = hg.g1_progression(5, 1, 3); Q = hg.g2_progression(6, 1, 3)
    P["inf"][1] = 1
    Q["inf"][2] = 1
    ml = ctx.miller_loop_batch(P, Q)
    one = np.zeros((2, 3, 2, 6), np.uint64); one[0, 0, 0] = _mont(1)
    assert (ml[1] == one).all() and (ml[2] == one).all()
    assert (ml[0] == orc.miller_loop(P[:1], Q[:1])).all()


def test_product_is_one_compare_two_pairings(ctx, orc):
    """CompareTwoPairings (pairing.go:140-147) as groups {(P1,Q1), (-P2,Q2)}: e(aG1, bG2) == e(abG1, G2)"""
    a, b = 0x1234567890abcdef1234, 0xfedcba0987654321
    ngroups = 37
    ps, qs, off, expect = [], [], [0], []
    for g in range(ngroups):
        ag, bg = a + g, b + 3 * g
        good = (g % 3) != 1
        rhs = ag * bg if good else ag * bg + 1
        ps += [hg.g1_mul(ag), hg.g1_neg(hg.g1_mul(rhs))]
        qs += [hg.g2_mul(bg), hg.g2_mul(1)]
        off.append(off[-1] + 2)
        expect.append(1 if good else 0)
    P = np.concatenate(ps); Q = np.concatenate(qs)
    ok = ctx.pairing_product_is_one(P, Q, off)
    assert ok.tolist() == expect
    assert orc.pairing_product_is_one(P, Q, off, threads=8).tolist() == expect


def test_product_ragged_groups(ctx, orc):
    """groups of different sizes incl. an empty one and a (P,Q),(−P,Q) cancellation of size 4"""
    P1 = hg.g1_progression(11, 7, 4); Q1 = hg.g2_progression(13, 5, 4)
    P = np.concatenate([P1[:1], hg.g1_neg(P1[:1]), P1[1:3], hg.g1_neg(P1[1:3]), P1[3:4]])
    Q = np.concatenate([Q1[:1], Q1[:1], Q1[1:3], Q1[1:3], Q1[3:4]])
    off = [0, 2, 2, 6, 7]
    ok = ctx.pairing_product_is_one(P, Q, off)
    assert ok.tolist() == [1, 1, 1, 0]      # the empty product is 1
    assert orc.pairing_product_is_one(P, Q, off).tolist() == [1, 1, 1, 0]


def test_product_large_groups_tree(ctx, orc):
    """VerifyAggregate-shaped products (g1pubs/bls.go:252-282): groups of hundreds of pairs are folded as trees
    (k_group_tree).  prod e((s+id)G1, (s'+id')G2) * e(-E G1, G2) == 1 with E = sum (s+id)(s'+id'), sizes 1..301 mixed
    with empty and two-pair groups; one group is broken on purpose; a small case is compared with the oracle."""
    s1, d1, s2, d2 = 0x1234567, 0x89, 0xabcdef1, 0x35
    sizes = [301, 0, 2, 64, 65, 1, 127, 2, 33]
    ps, qs, off, expect = [], [], [0], []
    for gi, m in enumerate(sizes):
        if m == 0:
            expect.append(1)
        elif m == 1:
            ps.append(hg.g1_mul(5)); qs.append(hg.g2_mul(7)); expect.append(0)
        else:
            k = m - 1
            P = hg.g1_progression(s1 + gi, d1, k); Q = hg.g2_progression(s2 + gi, d2, k)
            E = sum((s1 + gi + i * d1) * (s2 + gi + i * d2) for i in range(k)) % L.R_ORDER
            bad = gi == 6
            ps += [P, hg.g1_neg(hg.g1_mul(E + (1 if bad else 0)))]; qs += [Q, hg.g2_mul(1)]
            expect.append(0 if bad else 1)
        off.append(off[-1] + m)
    P = np.concatenate(ps); Q = np.concatenate(qs)
    assert ctx.pairing_product_is_one(P, Q, off).tolist() == expect
    lo, hi = off[7], off[9]                       # the last two groups (2 and 33 pairs) through the oracle as well
    sub = [o - lo for o in off[7:10]]
    assert orc.pairing_product_is_one(P[lo:hi], Q[lo:hi], sub, threads=8).tolist() == expect[7:9]
    assert ctx.pairing_product_is_one(P[lo:hi], Q[lo:hi], sub).tolist() == expect[7:9]


def test_bad_arguments(ctx):
    import ctypes
    from bls_b200 import capi
    assert ctx.lib.b381_pairing_batch(ctx._h, None, None, ctypes.c_size_t(4), None) == -1
    P = hg.g1_progression(1, 1, 2); Q = hg.g2_progression(1, 1, 2)
    with pytest.raises(capi.B381Error):
        ctx.pairing_product_is_one(P, Q, [0, 1])           # offsets do not cover npairs


def test_full_size_bilinearity_checksum(ctx, orc):
    """BASELINE config 2 at full size (2^16 pairings): prod_i e(a_i G1, b_i G2) == e(G1, G2)^(sum a_i b_i),
    checked as prod_i ML(P_i,Q_i) * ML(-S*G1, G2) -> FE == 1 on the device (one group of 2^16+1 pairs),
    plus every 1024th output bit-exact vs the oracle."""
    n = 1 << 16
    s, d, s2, d2 = 0xB2000002, 0x9E3779B97F4A7C15, 0x5EED5EED, 0xBF58476D1CE4E5B9
    P = hg.g1_progression(s, d, n); Q = hg.g2_progression(s2, d2, n)
    S = sum((s + i * d) * (s2 + i * d2) for i in range(n)) % L.R_ORDER
    Pn = np.concatenate([P, hg.g1_neg(hg.g1_mul(S))]); Qn = np.concatenate([Q, hg.g2_mul(1)])
    assert ctx.pairing_product_is_one(Pn, Qn, [0, n + 1]).tolist() == [1]
    Pn[-1] = hg.g1_neg(hg.g1_mul(S + 1))[0]
    assert ctx.pairing_product_is_one(Pn, Qn, [0, n + 1]).tolist() == [0]
    out = ctx.pairing_batch(P, Q)
    idx = np.arange(0, n, 1024)
    assert out[idx].tobytes() == orc.pairing_batch(P[idx], Q[idx], threads=8).tobytes()


def test_prepared_g2_points(ctx, orc):
    """b381_g2_prepare_batch == G2AffineToPrepared (g2.go:650-801: the oracle's 68 coefficient triples, bit for bit) and
    b381_miller_loop_prepared_batch == MillerLoop on a MillerLoopItem (pairing.go:4-7,16-75) == the fused Miller loop;
    several pairs share one prepared point through prep_idx; an infinite Q or P gives the factor 1"""
    n = 37
    Q = hg.g2_progression(0x5151, 0x77, 5)
    Q["inf"][4] = 1
    prep = ctx.g2_prepare_batch(Q)
    for i in range(4):
        assert (prep["coeffs"][i] == orc.g2_prepare(Q[i:i + 1])).all(), i
    assert prep["inf"].tolist() == [0, 0, 0, 0, 1]
    P = hg.g1_progression(0x1717, 0x3, n)
    P["inf"][9] = 1
    idx = (np.arange(n) % 5).astype(np.uint32)
    got = ctx.miller_loop_prepared_batch(P, prep, idx)
    want = ctx.miller_loop_batch(P, Q[idx])
    assert got.tobytes() == want.tobytes()
    for i in (0, 6, 13):
        assert (got[i] == orc.miller_loop(P[i:i + 1], Q[idx[i]:idx[i] + 1])).all(), i
    # without an index: pair i uses prepared point i
    got2 = ctx.miller_loop_prepared_batch(P[:5], prep)
    assert got2.tobytes() == ctx.miller_loop_batch(P[:5], Q).tobytes()
