"""The Python mirror of g1pubs / g2pubs (bls_b200/g1pubs.py, g2pubs.py, _pubs.py: every curve operation on the engine).

CPU part: the pure-Python oracle of the hashing / compression / key-derivation code (oracle/hostmath.py) against the
reference's own known-answer vectors (tests/golden/ref_kats.json "hash": hash_test.go:12-82,
g1pubs/bls_test.go:409-433) and against the C++ oracle; the key utilities of bls_b200/keys.py.
GPU part: the reference's API tests (g1pubs/bls_test.go:33-299, g2pubs/bls_test.go:33-213) re-run through the
engine -- BASELINE.json's first configuration (g2pubs Sign / Verify) included -- with the oracle's
CompareTwoPairings as the cross-check on the same points."""
import numpy as np
import pytest

from bls_b200 import hostgen as hg, keys as K, layout as L
from oracle import hostmath as hm


# ---- CPU: host mathematics -----------------------------------------------------------------------------
def test_hash_kats(kats):
    h = kats["hash"]
    msg = h["message"].encode()
    assert [hex(v) for v in hm.hash_g1(msg)] == h["hash_g1"]                                    # hash_test.go:12-26
    q = hm.hash_g2(msg)
    assert [hex(q[0][0]), hex(q[0][1]), hex(q[1][0]), hex(q[1][1])] == h["hash_g2_xc0_xc1_yc0_yc1"]   # :48-70
    assert hm.compress_g2(hm.hash_g2_with_domain(bytes(32), bytes(8))).hex() == h["hash_g2_with_domain_zero_compressed"]   # :72-82
    assert hex(hm.hash_secret_key(h["derive_secret_key_in"].encode())) == h["derive_secret_key_out"]  # g1pubs/bls_test.go:409-420
    assert hex(K.hash_secret_key(h["derive_secret_key_in"].encode())) == h["derive_secret_key_out"]


def test_hashed_points_are_in_the_groups():
    for m in (b"", b"a", b"Hello world! 16 characters 0"):
        p, q = hm.hash_g1(m), hm.hash_g2(m)
        assert (p[1] * p[1] - p[0] ** 3 - 4) % L.Q == 0 and hm.g1_in_subgroup(p)
        assert hm.g2_in_subgroup(q)


@pytest.mark.gpu
def test_invalid_pubkey_vectors(kats, pubs):
    """g1pubs/bls_test.go:422-433, g2pubs/bls_test.go:336-347: deserialisation must fail, not crash"""
    from bls_b200 import g1pubs, g2pubs
    with pytest.raises(ValueError):
        g1pubs.DeserializePublicKey(bytes.fromhex(kats["hash"]["invalid_g1_pubkey"]))
    with pytest.raises(ValueError):
        g2pubs.DeserializePublicKey(bytes.fromhex(kats["hash"]["invalid_g2_pubkey"]))


def test_compression_matches_oracle(orc):
    """CompressG1 / CompressG2 and back (g1.go:185-249, g2.go:219-289) on points with known discrete logs"""
    for k in (1, 2, 0x1234567, L.R_ORDER - 1):
        p, q = hm.g1_mul(hm.G1, k), hm.g2_mul(hm.G2, k)
        assert hm.compress_g1(p) == orc.g1.compress(hg.g1_points([p]))
        assert hm.compress_g2(q) == orc.g2.compress(hg.g2_points([q]))
        assert hm.decompress_g1(hm.compress_g1(p)) == (p, None)
        assert hm.decompress_g2(hm.compress_g2(q)) == (q, None)
    assert hm.compress_g1(None) == orc.g1.compress(hg.g1_points([None]))
    assert hm.decompress_g1(hm.compress_g1(None)) == (None, None)
    assert hm.decompress_g1(bytes(48))[1] == "unexpected compression mode"


def test_rand_key_matches_the_tests_reader(orc):
    """RandKey(NewXORShift(seed)) (g1_test.go:106-124 + crypto/rand.Int): same scalars as the oracle's reader"""
    for seed in (1, 2, 3, 20):
        r = K.XorShiftReader(seed)
        exp = orc.XorShift(seed).rand_fr(3)
        assert [K.rand_int(r, L.R_ORDER) for _ in range(3)] == [L.scalar_to_int(x) for x in exp]


# ---- GPU: the reference's API tests through the engine --------------------------------------------------
@pytest.fixture(scope="module")
def pubs():
    from bls_b200 import capi, g1pubs, g2pubs
    ctx = capi.Ctx(0)
    g1pubs.set_engine(ctx)
    yield g1pubs, g2pubs
    ctx.close()
    g1pubs.set_engine(None)


@pytest.mark.gpu
def test_g2pubs_sign_verify(pubs, orc):
    """BASELINE config 1 / g2pubs/bls_test.go:33-45 (xorshift seed 1 keys, the same messages)"""
    _, g2pubs = pubs
    r = K.XorShiftReader(1)
    for i in range(3):
        priv = g2pubs.RandKey(r)
        pub = g2pubs.PrivToPub(priv)
        msg = b"Hello world! 16 characters %d" % i
        sig = g2pubs.Sign(msg, priv)
        assert g2pubs.Verify(msg, pub, sig)
        assert not g2pubs.Verify(msg + b"!", pub, sig)
        # the oracle's CompareTwoPairings(sig, G2One, HashG1(m), pub) on the same points (pairing.go:140-147)
        assert orc.compare_two_pairings(orc.g1.to_proj(sig.s), orc.g2.to_proj(hg.g2_mul(1)),
                                        orc.g1.to_proj(hg.g1_points([hm.hash_g1(msg)])), orc.g2.to_proj(pub.p))
        # PrivToPub and Sign on the engine equal the oracle's scalar multiplications of the same points
        assert pub.p.tobytes() == hg.g2_mul(priv.f).tobytes()
        assert sig.s.tobytes() == hg.g1_points([hm.g1_mul(hm.hash_g1(msg), priv.f)]).tobytes()
        # serialisation round trips (g2pubs/bls_test.go:279-334)
        assert g2pubs.Verify(msg, g2pubs.DeserializePublicKey(pub.Serialize()), g2pubs.DeserializeSignature(sig.Serialize()))


@pytest.mark.gpu
def test_g1pubs_sign_verify_and_domain(pubs):
    """g1pubs/bls_test.go:33-45 and the WithDomain variants (g1pubs/bls.go:138-141, 171-174)"""
    g1pubs, _ = pubs
    r = K.XorShiftReader(1)
    priv = g1pubs.RandKey(r)
    pub = g1pubs.PrivToPub(priv)
    msg = b"Hello world! 16 characters 0"
    sig = g1pubs.Sign(msg, priv)
    assert g1pubs.Verify(msg, pub, sig) and not g1pubs.Verify(b"other", pub, sig)
    m32, dom = bytes(range(32)), bytes(8)
    sigd = g1pubs.SignWithDomain(m32, priv, dom)
    assert g1pubs.VerifyWithDomain(m32, pub, sigd, dom)
    assert not g1pubs.VerifyWithDomain(m32, pub, sigd, bytes([1]) + bytes(7))
    assert g1pubs.Verify(msg, g1pubs.DeserializePublicKey(pub.Serialize()), g1pubs.DeserializeSignature(sig.Serialize()))


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["g1pubs", "g2pubs"])
def test_aggregate_common_message(pubs, which):
    """SignVerifyAggregateCommonMessage + the missing-signature negative test (g1pubs/bls_test.go:47-129)"""
    mod = pubs[0] if which == "g1pubs" else pubs[1]
    r = K.XorShiftReader(2)
    msg = b">16 character identical message"
    keys = [mod.RandKey(r) for _ in range(6)]
    pubkeys = [mod.PrivToPub(k) for k in keys]
    sigs = [mod.Sign(msg, k) for k in keys]
    agg = mod.AggregateSignatures(sigs)
    assert agg.VerifyAggregateCommon(pubkeys, msg)
    assert not mod.AggregateSignatures(sigs[:-1]).VerifyAggregateCommon(pubkeys, msg)        # one signature missing
    assert not agg.VerifyAggregateCommon(pubkeys[:-1], msg)
    # incremental aggregation equals batch aggregation (Aggregate methods, bls.go:186-204)
    inc = mod.NewAggregateSignature()
    for s in sigs:
        inc.Aggregate(s)
    assert inc.Serialize() == agg.Serialize()
    ap = mod.NewAggregatePubkey()
    for p in pubkeys:
        ap.Aggregate(p)
    assert ap.Serialize() == mod.AggregatePublicKeys(pubkeys).Serialize()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["g1pubs", "g2pubs"])
def test_aggregate_distinct_messages(pubs, which):
    """SignVerifyAggregate + duplicate-message rejection (g1pubs/bls_test.go:131-218, bls.go:252-282)"""
    mod = pubs[0] if which == "g1pubs" else pubs[1]
    r = K.XorShiftReader(3)
    keys = [mod.RandKey(r) for _ in range(5)]
    pubkeys = [mod.PrivToPub(k) for k in keys]
    msgs = [b"Hello world! 16 characters %d" % i for i in range(5)]
    agg = mod.AggregateSignatures([mod.Sign(m, k) for m, k in zip(msgs, keys)])
    assert agg.VerifyAggregate(pubkeys, msgs)
    assert not agg.VerifyAggregate(pubkeys, msgs[:-1] + [b"tampered"])
    assert not agg.VerifyAggregate(pubkeys[:-1], msgs)                                          # length mismatch
    dup = msgs[:-1] + [msgs[0]]
    sig_dup = mod.AggregateSignatures([mod.Sign(m, k) for m, k in zip(dup, keys)])
    assert not sig_dup.VerifyAggregate(pubkeys, dup)                                             # duplicates are rejected
    assert not agg.VerifyAggregate(pubkeys, msgs[:-1] + [b""])                                   # quirk Q7: an empty message is rejected


@pytest.mark.gpu
def test_g1pubs_aggregate_with_domain(pubs):
    """VerifyAggregateCommonWithDomain / VerifyAggregateWithDomain (g1pubs/bls.go:294-311): the shape of
    verify_benchmark_test.go:33-85 at a small size"""
    g1pubs, _ = pubs
    r = K.XorShiftReader(5)
    dom = bytes([7]) + bytes(7)
    keys = [g1pubs.RandKey(r) for _ in range(4)]
    pubkeys = [g1pubs.PrivToPub(k) for k in keys]
    m = bytes(range(32))
    agg = g1pubs.AggregateSignatures([g1pubs.SignWithDomain(m, k, dom) for k in keys])
    assert agg.VerifyAggregateCommonWithDomain(pubkeys, m, dom)
    assert not agg.VerifyAggregateCommonWithDomain(pubkeys, bytes(32), dom)
    ms = [bytes([i]) * 32 for i in range(4)]
    agg2 = g1pubs.AggregateSignatures([g1pubs.SignWithDomain(mi, k, dom) for mi, k in zip(ms, keys)])
    assert agg2.VerifyAggregateWithDomain(pubkeys, ms, dom)
    assert not agg2.VerifyAggregateWithDomain(pubkeys, ms[::-1], dom)


@pytest.mark.gpu
def test_infinite_key_and_signature_do_not_verify(pubs):
    """Verify(m, Deserialize(infinity), Deserialize(infinity)) must be False for every message.  Both Miller
    pairs would be skipped (a pair with a point at infinity is the factor 1 in the engine; the reference panics there,
    pairing.go:17-26), and FE(1) == 1 would accept a zero-key forgery."""
    g1pubs, g2pubs = pubs
    inf48 = bytes([0xc0]) + bytes(47)
    inf96 = bytes([0xc0]) + bytes(95)
    pk1, sg1 = g1pubs.DeserializePublicKey(inf48), g1pubs.DeserializeSignature(inf96)
    assert not g1pubs.Verify(b"any message", pk1, sg1)
    assert not g1pubs.VerifyWithDomain(bytes(32), pk1, sg1, bytes(8))
    assert not sg1.VerifyAggregate([pk1], [b"m"])
    assert not sg1.VerifyAggregateCommon([pk1], b"m")
    assert not sg1.VerifyAggregateWithDomain([pk1], [bytes(32)], bytes(8))
    pk2, sg2 = g2pubs.DeserializePublicKey(inf96), g2pubs.DeserializeSignature(inf48)
    assert not g2pubs.Verify(b"any message", pk2, sg2)
    assert not sg2.VerifyAggregate([pk2], [b"m"])
    # a valid key with an infinite signature, and the reverse
    r = K.XorShiftReader(7)
    priv = g1pubs.RandKey(r); pub = g1pubs.PrivToPub(priv); sig = g1pubs.Sign(b"m", priv)
    assert g1pubs.Verify(b"m", pub, sig)
    assert not g1pubs.Verify(b"m", pub, sg1) and not g1pubs.Verify(b"m", pk1, sig)
