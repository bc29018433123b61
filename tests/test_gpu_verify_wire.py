"""b381_verify_with_domain_batch: g1pubs.VerifyWithDomain (g1pubs/bls.go:171-174) over wire-format triples, everything
(DeserializePublicKey/Signature, HashG2WithDomain, CompareTwoPairings) on the device.  Checked against the host
mirror of the reference API (bls_b200/g1pubs.py: host hashing + host deserialisation + engine pairing), against the
oracle's pairings, and by construction at batch scale (planted failures of every kind)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import codec_cases as cc

from bls_b200 import hostgen as hg, layout as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["auto", "vm", "thread", "duo"])
def ctx(request):
    from bls_b200 import capi
    return capi.Ctx(0, path=request.param)


def make_batch(ctx, n, seed):
    rng = np.random.RandomState(seed)
    sk = np.array([L.int_to_limbs(int.from_bytes(rng.bytes(31), "big") + 1, 4) for _ in range(n)], np.uint64)
    msgs = rng.randint(0, 256, (n, 32), dtype=np.uint8)
    domain = bytes(rng.randint(0, 256, 8, dtype=np.uint8))
    pubs = ctx.g1_compress_batch(ctx.g1_mul_batch(hg.g1_mul(1), sk))                  # PrivToPub + Serialize
    H = ctx.hash_g2_with_domain_batch(msgs, domain)
    sigs = ctx.g2_compress_batch(ctx.g2_mul_batch(H, sk))                             # SignWithDomain + Serialize
    return sk, msgs, domain, pubs, sigs


def test_against_host_mirror_and_oracle(ctx, orc):
    from bls_b200 import g1pubs
    g1pubs.set_engine(ctx)
    n = 12
    sk, msgs, domain, pubs, sigs = make_batch(ctx, n, 1)
    expect = np.ones(n, np.uint8)
    msgs[3, 0] ^= 1; expect[3] = 0                                   # wrong message
    sigs[5], sigs[6] = sigs[6].copy(), sigs[5].copy(); expect[5] = expect[6] = 0   # swapped signatures
    pubs[7] = pubs[8]; expect[7] = 0                                 # wrong key
    ok = ctx.verify_with_domain_batch(pubs, msgs, domain, sigs)
    assert ok.tolist() == expect.tolist()
    # the reference API, one call at a time (host hashing and deserialisation, engine pairing)
    for i in (0, 3, 5, 7, 9):
        pk = g1pubs.DeserializePublicKey(pubs[i].tobytes()); sg = g1pubs.DeserializeSignature(sigs[i].tobytes())
        assert g1pubs.VerifyWithDomain(msgs[i].tobytes(), pk, sg, domain) == bool(expect[i])
    # the oracle: e(G1One, sig) == e(pub, H(m)) as two pairings
    P, st = ctx.g1_decompress_batch(pubs.tobytes()); S, st2 = ctx.g2_decompress_batch(sigs.tobytes())
    H = ctx.hash_g2_with_domain_batch(msgs, domain)
    for i in (1, 3, 7):
        lhs = orc.pairing_batch(hg.g1_mul(1), S[i:i + 1]); rhs = orc.pairing_batch(P[i:i + 1], H[i:i + 1])
        assert (lhs.tobytes() == rhs.tobytes()) == bool(expect[i])


def test_bad_encodings_and_infinity_are_false(ctx):
    n = 8
    sk, msgs, domain, pubs, sigs = make_batch(ctx, n, 2)
    expect = np.ones(n, np.uint8)
    pubs[0] = np.frombuffer(cc.REF_INVALID_G1, np.uint8); expect[0] = 0          # g1pubs/bls_test.go:422-433
    sigs[1, 0] &= 0x7f; expect[1] = 0                                            # compression bit cleared
    inf48 = np.zeros(48, np.uint8); inf48[0] = 0xc0
    inf96 = np.zeros(96, np.uint8); inf96[0] = 0xc0
    pubs[2] = inf48; expect[2] = 0                                               # infinity key (the reference panics)
    sigs[3] = inf96; expect[3] = 0
    pubs[4] = inf48; sigs[4] = inf96; expect[4] = 0                              # e(G, O) == e(O, H) must not pass
    sigs[5, 95] ^= 1; expect[5] = 0                                              # x moved: off the curve or off the subgroup
    assert ctx.verify_with_domain_batch(pubs, msgs, domain, sigs).tolist() == expect.tolist()


def test_batch_scale_by_construction(ctx):
    n = 8192
    sk, msgs, domain, pubs, sigs = make_batch(ctx, n, 3)
    expect = np.ones(n, np.uint8)
    bad = np.arange(17, n, 97)
    msgs[bad, 31] ^= 0x80; expect[bad] = 0
    ok = ctx.verify_with_domain_batch(pubs, msgs, domain, sigs)
    assert ok.tolist() == expect.tolist()
    # per-item domains (stride 1) give the same verdicts when every item carries the same domain
    ok1 = ctx.verify_with_domain_batch(pubs[:512], msgs[:512], np.tile(np.frombuffer(domain, np.uint8), 512), sigs[:512])
    assert ok1.tolist() == expect[:512].tolist()


def make_plain_batch(ctx, n, seed, g2pubs=False):
    rng = np.random.RandomState(seed)
    sk = np.array([L.int_to_limbs(int.from_bytes(rng.bytes(31), "big") + 1, 4) for _ in range(n)], np.uint64)
    msgs = [rng.bytes(int(rng.randint(1, 100))) for _ in range(n)]
    if not g2pubs:      # g1pubs: keys in G1, signatures in G2 = sk * HashG2(m)
        pubs = ctx.g1_compress_batch(ctx.g1_mul_batch(hg.g1_mul(1), sk))
        sigs = ctx.g2_compress_batch(ctx.g2_mul_batch(ctx.hash_g2_batch(msgs), sk))
    else:               # g2pubs: keys in G2, signatures in G1 = sk * HashG1(m)
        pubs = ctx.g2_compress_batch(ctx.g2_mul_batch(hg.g2_mul(1), sk))
        sigs = ctx.g1_compress_batch(ctx.g1_mul_batch(ctx.hash_g1_batch(msgs), sk))
    return msgs, pubs, sigs


@pytest.mark.parametrize("pkg", ["g1pubs", "g2pubs"])
def test_plain_verify_against_host_mirror(ctx, pkg):
    """b381_g1pubs_verify_batch / b381_g2pubs_verify_batch against the host mirror of g1pubs.Verify (g1pubs/bls.go:165-168)
    and g2pubs.Verify (g2pubs/bls.go:159-162), plus planted failures at batch scale"""
    import importlib
    mod = importlib.import_module("bls_b200." + pkg)
    from bls_b200 import g1pubs
    g1pubs.set_engine(ctx)
    n = 1024
    msgs, pubs, sigs = make_plain_batch(ctx, n, 21, g2pubs=pkg == "g2pubs")
    expect = np.ones(n, np.uint8)
    for i in range(5, n, 37):
        msgs[i] = msgs[i] + b"!"; expect[i] = 0
    sigs[2], sigs[3] = sigs[3].copy(), sigs[2].copy(); expect[2] = expect[3] = 0
    pubs[8] = pubs[9]; expect[8] = 0
    ok = (ctx.g1pubs_verify_batch if pkg == "g1pubs" else ctx.g2pubs_verify_batch)(pubs, msgs, sigs)
    assert ok.tolist() == expect.tolist()
    for i in (0, 2, 5, 8, 11):
        pk = mod.DeserializePublicKey(pubs[i].tobytes()); sg = mod.DeserializeSignature(sigs[i].tobytes())
        assert mod.Verify(msgs[i], pk, sg) == bool(expect[i])


def test_rlc_batch_verification(ctx):
    """b381_verify_with_domain_rlc_batch: one boolean for the batch; accepts the all-valid batch and rejects as soon as one
    triple is wrong (wrong message, swapped signatures, undecodable key, infinity signature), for several weight draws"""
    n = 1024
    sk, msgs, domain, pubs, sigs = make_batch(ctx, n, 31)
    rng = np.random.RandomState(32)
    w = rng.randint(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, sigs, w) is True
    assert ctx.verify_with_domain_rlc_batch(pubs[:1], msgs[:1], domain, sigs[:1], w[:1]) is True
    m2 = msgs.copy(); m2[777, 5] ^= 4
    assert ctx.verify_with_domain_rlc_batch(pubs, m2, domain, sigs, w) is False
    s2 = sigs.copy(); s2[10], s2[11] = sigs[11], sigs[10]
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, s2, w) is False
    p2 = pubs.copy(); p2[3] = np.frombuffer(cc.REF_INVALID_G1, np.uint8)
    assert ctx.verify_with_domain_rlc_batch(p2, msgs, domain, sigs, w) is False
    s3 = sigs.copy(); s3[0] = 0; s3[0, 0] = 0xc0
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, s3, w) is False
    # the per-item path agrees on which item is wrong
    ok = ctx.verify_with_domain_batch(pubs, m2, domain, sigs)
    assert ok.sum() == n - 1 and ok[777] == 0


def test_rlc_weights_zero_or_too_wide_are_rejected(ctx):
    """a zero weight would silently drop its triple (r pk = infinity is the factor 1), so a batch whose only bad
    signature carries weight 0 would pass; weights wider than the declared bit length would be truncated.  Both make the check
    false; weights drawn by the wrapper (None) are non-zero 64-bit values."""
    _, msgs, domain, pubs, sigs = make_batch(ctx, 8, 31)
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, sigs) is True                 # wrapper-generated weights
    w = np.arange(1, 9, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, sigs, w) is True
    s2 = sigs.copy(); s2[3] = sigs[4]                                                          # a wrong signature ...
    w0 = w.copy(); w0[3] = 0                                                                   # ... hidden behind a zero weight
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, s2, w) is False
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, s2, w0) is False
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, sigs, w0) is False             # zero weight on a valid batch too
    ctx.set_rlc_weight_bits(32)                                                                # declared 32 bits, weights are 64 bits wide
    r = np.zeros((8, 4), np.uint64); r[:, 0] = w
    import ctypes
    from bls_b200.capi import _hp
    ok = np.ones(1, np.uint8)
    p = np.ascontiguousarray(pubs, np.uint8).reshape(-1); m = np.ascontiguousarray(msgs, np.uint8).reshape(-1)
    sg = np.ascontiguousarray(sigs, np.uint8).reshape(-1); d = np.frombuffer(bytes(domain), np.uint8).copy()
    ctx.call("b381_verify_with_domain_rlc_batch", _hp(p), _hp(m), _hp(d), ctypes.c_size_t(0), _hp(sg), _hp(r), ctypes.c_size_t(8), _hp(ok))
    assert ok[0] == 0
    ctx.set_rlc_weight_bits(255)


def test_rlc_partials_combine_like_the_unsharded_check(ctx):
    """the multi-GPU form on one device: three 'ranks' form partial Miller products of their tiles; the product of the
    partials passes the final exponentiation iff the unsharded batch does (b381_verify_rlc_partial_dev +
    b381_fp12_product_final_exp_is_one_dev, the exchange of bls_b200/dist.py::verify_rlc_sharded)"""
    n = 96
    sk, msgs, domain, pubs, sigs = make_batch(ctx, n, 41)
    P, _ = ctx.g1_decompress_batch(pubs.tobytes()); S, _ = ctx.g2_decompress_batch(sigs.tobytes())
    H = ctx.hash_g2_with_domain_batch(msgs, domain)
    w = np.random.RandomState(42).randint(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
    cuts = [0, 17, 64, 96]

    def sharded(Hx):
        parts, valid = [], []
        for a, b in zip(cuts, cuts[1:]):
            p, v = ctx.verify_rlc_partial(P[a:b], Hx[a:b], S[a:b], w[a:b])
            parts.append(p); valid.append(v)
        return all(valid) and ctx.fp12_product_final_exp_is_one(np.concatenate(parts))
    assert sharded(H) is True
    H2 = H.copy(); H2[70] = H[71]
    assert sharded(H2) is False
    # the partial of an empty tile is the neutral element
    p0, v0 = ctx.verify_rlc_partial(P[:0], H[:0], S[:0], w[:0])
    assert v0 == 1 and ctx.fp12_product_final_exp_is_one(p0)
    # b381_miller_product_dev against the product of the oracle-checked Miller loops
    # b381_miller_product_dev: e(aG1, bG2) * e(-(ab)G1, G2) has Miller product whose final exponentiation is one
    a, b = 0x1234, 0x5678
    Pp = np.concatenate([hg.g1_mul(a), hg.g1_neg(hg.g1_mul(a * b))]); Qq = np.concatenate([hg.g2_mul(b), hg.g2_mul(1)])
    assert ctx.fp12_product_final_exp_is_one(ctx.miller_product(Pp, Qq))
    assert not ctx.fp12_product_final_exp_is_one(ctx.miller_product(Pp[:1], Qq[:1]))


def test_empty_and_single_item_batches(ctx):
    """ragged ends of every new entry point: n = 0 and n = 1"""
    assert ctx.verify_with_domain_batch(np.zeros((0, 48), np.uint8), np.zeros((0, 32), np.uint8), bytes(8), np.zeros((0, 96), np.uint8)).size == 0
    assert ctx.g1pubs_verify_batch(np.zeros((0, 48), np.uint8), [], np.zeros((0, 96), np.uint8)).size == 0
    assert ctx.g2pubs_verify_batch(np.zeros((0, 96), np.uint8), [], np.zeros((0, 48), np.uint8)).size == 0
    assert ctx.hash_g1_batch([]).size == 0 and ctx.hash_g2_batch([]).size == 0
    assert not ctx.g2_msm(np.zeros(0, dtype=L.G2_AFFINE), np.zeros((0, 4), np.uint64))["z"].any()
    sk, msgs, domain, pubs, sigs = make_batch(ctx, 1, 77)
    assert ctx.verify_with_domain_batch(pubs, msgs, domain, sigs).tolist() == [1]
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, sigs, np.array([12345], np.uint64)) is True
    msgs[0, 0] ^= 1
    assert ctx.verify_with_domain_batch(pubs, msgs, domain, sigs).tolist() == [0]
    assert ctx.verify_with_domain_rlc_batch(pubs, msgs, domain, sigs, np.array([12345], np.uint64)) is False
    m1, p1, s1 = make_plain_batch(ctx, 1, 78)
    assert ctx.g1pubs_verify_batch(p1, m1, s1).tolist() == [1]
    assert ctx.g1pubs_verify_batch(p1, [b""], s1).tolist() == [0]          # the empty message is a valid input to HashG2
