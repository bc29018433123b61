"""Parity of the aggregation kernels (point sums, Pippenger MSM, bucket-sharded MSM + fold, the
VerifyAggregateCommon batch) against the CPU oracle, through the C ABI.  Results are compared on
canonical observables: affine coordinates (Montgomery limbs) and booleans, bit-exact."""
import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L

pytestmark = pytest.mark.gpu


# the verification tests below run on the VM kernels ("auto" at these sizes), on the one-pairing-per-thread kernels incl. the
# shared-accumulator k_miller_loop2, and on the two-lane kernels (k_duo_miller_loop<2>): b381_set_kernel_path
@pytest.fixture(scope="module", params=["auto", "vm", "thread", "duo"])
def ctx(request):
    from bls_b200 import capi
    c = capi.Ctx(0, path=request.param)
    yield c
    c.close()


def _same_point(orc, grp, got_jac, exp_jac):
    a = grp.to_affine(got_jac); b = grp.to_affine(exp_jac)
    return a.tobytes() == b.tobytes()


def _normalised(orc, jac, fp2=False):
    """engine sums come back with z == 1 (or the canonical zero)"""
    one = L.fp_from_int(1)
    z = jac["z"][0]
    if fp2:
        return (z[0] == one).all() and not z[1].any() or not z.any()
    return (z == one).all() or not z.any()


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 1000, 5000])
def test_g1_sum_vs_reference_fold(ctx, orc, n):
    """AggregatePublicKeys (g1pubs/bls.go:192-198): left fold of G1Projective.Add"""
    P = hg.g1_progression(0xA11CE + n, 0x1D, n) if n else np.zeros(0, dtype=L.G1_AFFINE)
    got = ctx.g1_sum(P)
    assert _normalised(orc, got)
    assert _same_point(orc, orc.g1, got, orc.g1.sum_affine(P))


def test_g1_sum_special_cases(ctx, orc):
    """equal points (doubling branch, g1.go:419-421), P + (-P), infinity inputs (g1.go:401-406)"""
    P = hg.g1_progression(5, 3, 6)
    dup = np.concatenate([P[:1], P[:1]])                       # P + P
    assert _same_point(orc, orc.g1, ctx.g1_sum(dup), orc.g1.sum_affine(dup))
    canc = np.concatenate([P[:3], hg.g1_neg(P[:3])])           # sums to zero
    got = ctx.g1_sum(canc)
    assert not got["z"].any() and _same_point(orc, orc.g1, got, orc.g1.sum_affine(canc))
    withinf = P.copy(); withinf["inf"][2] = 1; withinf["inf"][5] = 1
    assert _same_point(orc, orc.g1, ctx.g1_sum(withinf), orc.g1.sum_affine(withinf))
    many = np.resize(P[:1], 300)                               # 300 * P: every tree level doubles
    assert _same_point(orc, orc.g1, ctx.g1_sum(many), orc.g1.sum_affine(many))


@pytest.mark.parametrize("n", [0, 1, 3, 64, 65, 700])
def test_g2_sum_vs_reference_fold(ctx, orc, n):
    """g1pubs.AggregateSignatures (g1pubs/bls.go:177-183): fold of G2Projective.Add"""
    Q = hg.g2_progression(0xB0B + n, 0x11, n) if n else np.zeros(0, dtype=L.G2_AFFINE)
    got = ctx.g2_sum(Q)
    assert _normalised(orc, got, fp2=True)
    assert _same_point(orc, orc.g2, got, orc.g2.sum_affine(Q))
    if n >= 3:
        dup = np.concatenate([Q[:2], Q[:2], Q[2:3]])
        assert _same_point(orc, orc.g2, ctx.g2_sum(dup), orc.g2.sum_affine(dup))


@pytest.mark.parametrize("n", [1, 2, 33, 500, 3000])
def test_g1_msm_vs_double_and_add(ctx, orc, n):
    """sum k_i P_i == fold of G1Affine.MulFR (g1.go:80-90) with G1Projective.Add"""
    P = hg.g1_progression(0x5EED + n, 0x77, n)
    K, _ = hg.splitmix_scalars(n, n)
    got = ctx.g1_msm(P, K)
    assert _normalised(orc, got)
    assert _same_point(orc, orc.g1, got, orc.g1_msm_naive(P, K, threads=8))


def test_g1_msm_degenerate_scalars(ctx, orc):
    """all-ones scalars (== AggregatePublicKeys: every point in one bucket), zeros, r-1, repeated points"""
    n = 2000
    P = hg.g1_progression(9, 4, n)
    ones = np.zeros((n, 4), np.uint64); ones[:, 0] = 1
    assert _same_point(orc, orc.g1, ctx.g1_msm(P, ones), orc.g1.sum_affine(P))
    zeros = np.zeros((n, 4), np.uint64)
    assert not ctx.g1_msm(P, zeros)["z"].any()
    K, vals = hg.splitmix_scalars(7, 40)
    K[0] = L.scalar_from_int(L.R_ORDER - 1); K[1] = L.scalar_from_int(0); K[2] = L.scalar_from_int((1 << 255) - 1 - 3)
    Pm = np.resize(P[:3], 40)                                  # the same 3 points over and over
    Pm["inf"][5] = 1
    assert _same_point(orc, orc.g1, ctx.g1_msm(Pm, K), orc.g1_msm_naive(Pm, K, threads=8))
    assert not ctx.g1_msm(np.zeros(0, dtype=L.G1_AFFINE), np.zeros((0, 4), np.uint64))["z"].any()


def test_g1_msm_closed_form_large(ctx, orc):
    """2^17 points with known discrete logs: sum k_i (s + i d) G1 == (sum k_i (s + i d) mod r) G1"""
    n = 1 << 17
    s, d = 0xB2000003, 0x9E3779B97F4A7C15
    P = hg.g1_progression(s, d, n)
    K, vals = hg.splitmix_scalars(42, n)
    S = sum(k * (s + i * d) for i, k in enumerate(vals)) % L.R_ORDER
    got = orc.g1.to_affine(ctx.g1_msm(P, K))
    assert got.tobytes() == hg.g1_mul(S).tobytes()


def test_full_size_configs_closed_form(ctx, orc):
    """BASELINE.json's full sizes through closed forms: (config 4) a 2^22-point G1 MSM with canonical 255-bit scalars and
    (config 3) the aggregate of 2^20 public keys.  2^14 distinct points with known discrete logs are tiled (the work
    does not depend on the values) and 2^12 distinct scalars repeat, so the expected point is one big-integer sum."""
    m, s, d = 1 << 14, 0xA66, 0x51
    base = hg.g1_progression(s, d, m)
    K, kv = hg.splitmix_scalars(99, 1 << 12)
    n = 1 << 22
    # sum_i k_(i mod 4096) * (s + (i mod m) d): i mod 4096 and i mod m are both determined by i mod m (4096 | m)
    per_tile = sum(kv[i % 4096] * (s + i * d) for i in range(m)) % L.R_ORDER
    S = per_tile * (n // m) % L.R_ORDER
    got = orc.g1.to_affine(ctx.g1_msm(np.resize(base, n), np.resize(K, (n, 4))))
    assert got.tobytes() == hg.g1_mul(S).tobytes()
    n = 1 << 20
    sk_sum = sum(s + i * d for i in range(m)) * (n // m) % L.R_ORDER
    agg = orc.g1.to_affine(ctx.g1_sum(np.resize(base, n)))
    assert agg.tobytes() == hg.g1_mul(sk_sum).tobytes()


def test_g2_msm(ctx, orc):
    """b381_g2_msm against the fold of G2Affine.MulFR (g2.go:92-102) on a small case, the closed form
    sum k_i (s + i d) G2 on 2^14 points, 64-bit weights (the random-linear-combination shape) and degenerate inputs"""
    n = 1 << 14
    s, d = 0x5EED5, 0x9E3779B9
    P = hg.g2_progression(s, d, n)
    K, vals = hg.splitmix_scalars(7, n)
    S = sum(k * (s + i * d) for i, k in enumerate(vals)) % L.R_ORDER
    assert orc.g2.to_affine(ctx.g2_msm(P, K)).tobytes() == hg.g2_mul(S).tobytes()
    m = 40
    exp = orc.g2.sum_proj(orc.g2.mul_fr(P[:m], K[:m], threads=8))
    assert orc.g2.to_affine(ctx.g2_msm(P[:m], K[:m])).tobytes() == orc.g2.to_affine(exp).tobytes()
    K64 = K.copy(); K64[:, 1:] = 0
    S64 = sum(int(K64[i, 0]) * (s + i * d) for i in range(n)) % L.R_ORDER
    assert orc.g2.to_affine(ctx.g2_msm(P, K64)).tobytes() == hg.g2_mul(S64).tobytes()
    Z = np.zeros_like(K[:8])
    assert not ctx.g2_msm(P[:8], Z)["z"].any()                       # all-zero scalars: the point at infinity
    same = np.repeat(K[:1], 3000, axis=0)                            # every point in one bucket per window
    Ssame = vals[0] * sum(s + i * d for i in range(3000)) % L.R_ORDER
    assert orc.g2.to_affine(ctx.g2_msm(P[:3000], same)).tobytes() == hg.g2_mul(Ssame).tobytes()


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_g1_msm_bucket_sharded(ctx, orc, nranks):
    """BASELINE config 4 semantics on one GPU: every rank's window shard, then the fold of the partials"""
    n = 1500
    P = hg.g1_progression(0xC0FFEE, 0x31, n)
    K, _ = hg.splitmix_scalars(nranks, n)
    parts = np.concatenate([ctx.g1_msm_shard(P, K, r, nranks) for r in range(nranks)])
    got = ctx.g1_fold(parts)
    assert _normalised(orc, got)
    assert _same_point(orc, orc.g1, got, orc.g1_msm_naive(P, K, threads=8))
    assert got.tobytes() == ctx.g1_msm(P, K).tobytes()


@pytest.mark.parametrize("nranks", [1, 2])
def test_g1_msm_exceptional_cases_in_the_reductions(ctx, orc, nranks):
    """bucket sums that are EQUAL (the running sums of the bucket reduction must double, g1.go:506-509) or OPPOSITE (they pass
    through infinity), in every window a 255-bit scalar touches: the lane-cooperative additions of the latency-bound reduction
    (small and bucket-sharded MSMs) and the one-thread-per-segment form both take these branches"""
    P = hg.g1_mul(0xABCDEF12345); Q = hg.g1_mul(0x1234567)
    pts = np.concatenate([P, P, hg.g1_neg(P), Q, Q, Q, hg.g1_neg(Q), P])
    rng = np.random.RandomState(3)
    ks = [5, 3, 4, 7, 6, 2, 6, 1]                                       # window 0: B5 = P, B3 = P, B4 = -P, B7 = B6... = Q, -Q + Q
    top = [int(x) for x in rng.randint(1, 1 << 30, size=8)]
    vals = []
    for i, k in enumerate(ks):                                          # the same small digits again in a high window + noise between
        vals.append(k | (top[i % 2] << 64) | (k << 200))
    K = np.concatenate([L.scalar_from_int(v % L.R_ORDER).reshape(1, 4) for v in vals])
    want = orc.g1_msm_naive(pts, K, threads=2)
    parts = np.concatenate([ctx.g1_msm_shard(pts, K, r, nranks) for r in range(nranks)])
    got = ctx.g1_fold(parts)
    assert _same_point(orc, orc.g1, got, want)
    assert got.tobytes() == ctx.g1_msm(pts, K).tobytes()
    # many copies: every bucket of a window holds the same point sum
    n = 3000
    same = np.resize(P, n)
    K2, _ = hg.splitmix_scalars(11, n)
    parts = np.concatenate([ctx.g1_msm_shard(same, K2, r, nranks) for r in range(nranks)])
    assert _same_point(orc, orc.g1, ctx.g1_fold(parts), orc.g1_msm_naive(same, K2, threads=8))


def _attestations(orc, nreg, natt, committee, nmsg, seed):
    """registry of pk_i = sk_i G1 with sk_i = s + i d; messages as points H_j = h_j G2 (stand-ins for
    HashG2WithDomain outputs, which the Go host computes); sig_a = (sum of the committee's sk) * H_j"""
    s, d = 0x1000 + seed, 0x2B
    reg = hg.g1_progression(s, d, nreg)
    hs = [0x77 + 5 * j for j in range(nmsg)]
    H = np.concatenate([hg.g2_mul(h) for h in hs])
    rng = np.random.RandomState(seed)
    key_idx, key_off, sigs, msg_idx, expect = [], [0], [], [], []
    for a in range(natt):
        m = committee if a % 5 else max(1, committee // 3)       # ragged committees
        ks = rng.randint(0, nreg, size=m)
        j = int(rng.randint(0, nmsg))
        sk = sum(s + int(i) * d for i in ks) % L.R_ORDER
        bad = a % 4 == 3
        sigs.append(hg.g2_mul((sk + (1 if bad else 0)) * hs[j]))
        key_idx += [int(i) for i in ks]; key_off.append(len(key_idx)); msg_idx.append(j); expect.append(0 if bad else 1)
    return reg, np.array(key_idx, np.uint32), np.array(key_off, np.uint32), np.concatenate(sigs), H, np.array(msg_idx, np.uint32), expect


def test_verify_aggregate_common_batch(ctx, orc):
    """VerifyAggregateCommon (g1pubs/bls.go:287-290) per attestation, incl. corrupted signatures"""
    reg, kidx, koff, sig, H, midx, expect = _attestations(orc, 64, 21, 9, 3, 1)
    ok = ctx.verify_aggregate_common_batch(reg, kidx, koff, sig, H, midx)
    assert ok.tolist() == expect
    # the oracle's own path: AggregatePublicKeys fold + CompareTwoPairings(G1One, sig, pk, H)
    g1one = orc.g1.to_proj(orc.g1_generator())
    for a in range(len(expect)):
        pk = orc.g1.sum_affine(reg[kidx[koff[a]:koff[a + 1]]])
        good = orc.compare_two_pairings(g1one, orc.g2.to_proj(sig[a:a + 1]), pk, orc.g2.to_proj(H[midx[a]:midx[a] + 1]))
        assert int(good) == expect[a] == int(ok[a])


def test_attestation_batch_as_one_random_linear_combination(ctx, orc):
    """b381_verify_aggregate_common_rlc_dev: one boolean for the batch == AND of the per-attestation VerifyAggregateCommon
    verdicts (which are checked against the oracle above); ragged committees, several messages, a message nobody signs,
    every single corrupted position, bad weights, bad indices"""
    reg, kidx, koff, sig, H, midx, expect = _attestations(orc, 64, 24, 9, 3, 5)
    good = [a for a, e in enumerate(expect) if e]                      # the construction corrupts every fourth signature
    assert len(good) < len(expect)
    assert ctx.verify_aggregate_common_rlc(reg, kidx, koff, sig, H, midx) is False
    assert ctx.verify_aggregate_common_batch(reg, kidx, koff, sig, H, midx).tolist() == expect

    def subset(idx):
        ki = np.concatenate([kidx[koff[a]:koff[a + 1]] for a in idx]).astype(np.uint32)
        ko = np.concatenate([[0], np.cumsum([koff[a + 1] - koff[a] for a in idx])]).astype(np.uint32)
        return ki, ko, sig[idx], midx[idx]
    ki, ko, sg, mi = subset(good)
    H4 = np.concatenate([H, hg.g2_mul(0x4242)])                        # a fourth message that no attestation refers to
    assert ctx.verify_aggregate_common_rlc(reg, ki, ko, sg, H4, mi) is True
    assert ctx.verify_aggregate_common_rlc(reg, ki, ko, sg, H4, mi, weights=np.arange(1, len(good) + 1)) is True
    # one bad attestation anywhere makes the batch false
    bad_a = next(a for a, e in enumerate(expect) if not e)
    for pos in (0, len(good) // 2, len(good)):
        idx = good[:pos] + [bad_a] + good[pos:]
        ki2, ko2, sg2, mi2 = subset(idx)
        assert ctx.verify_aggregate_common_rlc(reg, ki2, ko2, sg2, H4, mi2) is False, pos
    # two corruptions that cancel without weights: sig_0 + D and sig_1 - D for the same message are caught by random weights
    same = [a for a in good if midx[a] == midx[good[0]]][:2]
    assert len(same) == 2
    ki3, ko3, sg3, mi3 = subset(same)
    s0, d0, h = 0x1000 + 5, 0x2B, 0x77 + 5 * int(midx[same[0]])       # the discrete logs _attestations(seed = 5) used
    sks = [sum(s0 + int(i) * d0 for i in kidx[koff[a]:koff[a + 1]]) for a in same]
    assert hg.g2_mul(sks[0] * h % L.R_ORDER).tobytes() == sg3[0:1].tobytes()
    sg3 = np.concatenate([hg.g2_mul((sks[0] * h + 0x999) % L.R_ORDER), hg.g2_mul((sks[1] * h - 0x999) % L.R_ORDER)])
    assert ctx.verify_aggregate_common_rlc(reg, ki3, ko3, sg3, H4, mi3, weights=[1, 1]) is True       # the forgery equal weights admit
    assert ctx.verify_aggregate_common_rlc(reg, ki3, ko3, sg3, H4, mi3) is False                      # ... and random ones reject
    # fail closed: zero weight, over-wide weight (bits = 64 is set by the wrapper), message / key index outside the tables, empty batch
    w = np.arange(1, len(good) + 1).astype(np.uint64); w[3] = 0
    assert ctx.verify_aggregate_common_rlc(reg, ki, ko, sg, H4, mi, weights=w) is False
    mi_bad = mi.copy(); mi_bad[2] = 77
    assert ctx.verify_aggregate_common_rlc(reg, ki, ko, sg, H4, mi_bad) is False
    ki_bad = ki.copy(); ki_bad[5] = 64
    assert ctx.verify_aggregate_common_rlc(reg, ki_bad, ko, sg, H4, mi) is False
    sg_inf = sg.copy(); sg_inf["inf"][1] = 1
    assert ctx.verify_aggregate_common_rlc(reg, ki, ko, sg_inf, H4, mi) is False
    assert ctx.verify_aggregate_common_rlc(reg, np.zeros(0, np.uint32), np.zeros(1, np.uint32), sg[:0], H4, mi[:0]) is True


def test_verify_aggregate_empty_committee_and_batch(ctx, orc):
    """an empty committee aggregates to the zero key (the reference panics in MillerLoop; the
    engine treats the infinity pair as the factor 1, so the check reduces to e(G1, sig) == 1: false)"""
    reg, kidx, koff, sig, H, midx, expect = _attestations(orc, 16, 3, 4, 2, 2)
    koff2 = koff.copy(); koff2[1:] = koff2[1]                   # attestations 1, 2 have no keys
    ok = ctx.verify_aggregate_common_batch(reg, kidx, koff2, sig, H, midx)
    assert ok.tolist() == [expect[0], 0, 0]
    assert ctx.verify_aggregate_common_batch(reg, kidx[:0], np.zeros(1, np.uint32), sig[:0], H, midx[:0]).size == 0


def test_attestation_batch_fails_closed(ctx, orc):
    """an attestation must not verify by degenerating.  With an infinite aggregate signature the check of an EMPTY
    committee (or of keys that cancel) would reduce to 1 == 1, because the engine's Miller loop treats a pair with a point at
    infinity as the factor 1 where the reference panics (pairing.go:17-26).  Such attestations, and ones whose key / message
    indices leave the tables, are reported false."""
    reg, kidx, koff, sig, H, midx, expect = _attestations(orc, 16, 5, 4, 2, 9)
    base = ctx.verify_aggregate_common_batch(reg, kidx, koff, sig, H, midx).tolist()
    assert base == expect
    # (1) empty committee + infinite signature
    koff1 = koff.copy(); koff1[1:] = koff1[1]
    sig1 = sig.copy(); sig1["inf"][1] = 1
    assert ctx.verify_aggregate_common_batch(reg, kidx, koff1, sig1, H, midx).tolist()[1] == 0
    # (2) a committee whose keys cancel (P and -P) + infinite signature
    reg2 = np.concatenate([reg, hg.g1_neg(reg[:1])])
    assert koff[3] - koff[2] == 4
    kidx2 = kidx.copy(); kidx2[koff[2]:koff[3]] = [0, reg.size, 0, reg.size]
    sig2 = sig.copy(); sig2["inf"][2] = 1
    assert ctx.verify_aggregate_common_batch(reg2, kidx2, koff, sig2, H, midx).tolist()[2] == 0
    # (3) an infinite signature alone
    sig3 = sig.copy(); sig3["inf"][0] = 1
    assert ctx.verify_aggregate_common_batch(reg, kidx, koff, sig3, H, midx).tolist()[0] == 0
    # (4) indices outside the registry / the message table
    kidx4 = kidx.copy(); kidx4[koff[3]] = reg.size + 7
    midx4 = midx.copy(); midx4[4] = H.size
    got = ctx.verify_aggregate_common_batch(reg, kidx4, koff, sig, H, midx4).tolist()
    assert got[3] == 0 and got[4] == 0 and got[:3] == expect[:3]


def test_attestations_from_wire(ctx, orc):
    """b381_verify_aggregate_common_with_domain_batch_dev: VerifyAggregateCommonWithDomain (g1pubs/bls.go:294-297) with
    compressed aggregate signatures and 32-byte message hashes; verdicts by construction (known secret keys), an
    undecodable signature and an infinite one are false, and the pre-hashed entry point agrees"""
    m, s0, d0 = 64, 0xA66, 0x51
    keys = hg.g1_progression(s0, d0, m)
    rng = np.random.RandomState(5)
    natt, comm, nmsg = 24, 9, 3
    msgs = rng.randint(0, 256, (nmsg, 32), dtype=np.uint8)
    domain = bytes(range(8))
    H = ctx.hash_g2_with_domain_batch(msgs, domain)
    tk = rng.randint(0, m, size=(natt, comm))
    midx = (np.arange(natt) % nmsg).astype(np.uint32)
    sk = [sum(s0 + int(i) * d0 for i in tk[a]) % L.R_ORDER for a in range(natt)]
    expect = np.ones(natt, np.uint8)
    sk[4] += 1; expect[4] = 0                                                # wrong aggregate signature
    K = np.array([L.int_to_limbs(k, 4) for k in sk], np.uint64)
    sig_aff = ctx.g2_mul_batch(H[midx], K)                                   # sum of the members' signatures on H(m)
    sigs = ctx.g2_compress_batch(sig_aff)
    sigs[7, 0] &= 0x7f; expect[7] = 0                                        # compression bit cleared
    sigs[9] = 0; sigs[9, 0] = 0xc0; expect[9] = 0                            # infinity
    kidx = tk.reshape(-1).astype(np.uint32); koff = (np.arange(natt + 1) * comm).astype(np.uint32)
    ok = ctx.verify_aggregate_common_with_domain_batch(keys, kidx, koff, sigs, msgs, domain, midx)
    assert ok.tolist() == expect.tolist()
    good = [a for a in range(natt) if a not in (7, 9)]
    ok2 = ctx.verify_aggregate_common_batch(keys, kidx, koff, sig_aff, H, midx)
    assert ok2[good].tolist() == expect[good].tolist()
