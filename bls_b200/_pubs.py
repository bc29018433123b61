"""Shared implementation of the Python mirrors of g1pubs / g2pubs on top of the engine.  EVERY piece of curve work --
hashing to the curve, (de)serialisation with subgroup checks, scalar multiplication for key generation and signing,
aggregation, pairing checks -- is a call into libb381.so (include/b381.h); the only arithmetic left on the host is
the sign flip Q - y of a canonical coordinate (NegAssign in the Go shim).  Points are single-element numpy arrays of the ABI PODs (layout.G1_AFFINE / G2_AFFINE)."""
import numpy as np

from . import keys as K, layout as L

_ctx = None
DECODE_ERRORS = {1: "unexpected compression mode", 2: "unexpected information in compressed infinity",
                 3: "point not on curve", 4: "not in correct subgroup"}       # g1.go:120,192,204,213; g2.go:158,226,234,242


def engine():
    global _ctx
    if _ctx is None:
        from . import capi
        _ctx = capi.Ctx(0)
    return _ctx


def set_engine(ctx):
    global _ctx
    _ctx = ctx


class Group:
    """the engine entry points of one group (G1 or G2)"""

    def __init__(self, which):
        self.which = which
        self.dtype = L.G1_AFFINE if which == 1 else L.G2_AFFINE
        self.nbytes = 48 if which == 1 else 96

    def f(self, name):
        return getattr(engine(), "g%d_%s" % (self.which, name))

    def zero(self):
        z = np.zeros(1, dtype=self.dtype)
        z["inf"] = 1
        y = z["y"].reshape(-1, 6)
        y[0] = L.fp_from_int(1)
        return z

    def generator(self):
        """G1One / G2One (g1.go:25-32, g2.go:26-43)"""
        return (_G1_GEN if self.which == 1 else _G2_GEN).copy()

    def compress(self, p):
        return self.f("compress_batch")(p)[0].tobytes()

    def decompress(self, b):
        pts, st = self.f("decompress_batch")(bytes(b), check_subgroup=True)
        if st[0]:
            raise ValueError(DECODE_ERRORS[int(st[0])])
        return pts

    def mul(self, p, k):
        """k * p for a point of the group (the generator or a hash to the curve): PrivToPub / Sign"""
        return self.f("mul_subgroup_batch")(p, np.array([L.int_to_limbs(k % L.R_ORDER, 4)], np.uint64))

    def sum(self, pts):
        j = self.f("sum")(np.concatenate(pts)) if pts else None
        if j is None or not j["z"].any():
            return self.zero()
        out = np.zeros(1, dtype=self.dtype)
        out["x"] = j["x"]; out["y"] = j["y"]                      # the engine returns z = 1
        return out

    def neg(self, p):
        """-P: multiply by r - 1 would do; negation is a host-side limb subtraction Q - y on a canonical value"""
        out = p.copy()
        if not out["inf"][0]:
            y = out["y"].reshape(-1, 6)
            for i in range(y.shape[0]):
                v = L.limbs_to_int(y[i])
                y[i] = np.array(L.int_to_limbs((L.Q - v) % L.Q), dtype=np.uint64)
        return out

    def hash(self, msg):
        return (engine().hash_g1_batch if self.which == 1 else engine().hash_g2_batch)([bytes(msg)])


def _gen_pod(dtype, coords):
    p = np.zeros(1, dtype=dtype)
    x = p["x"].reshape(-1, 6); y = p["y"].reshape(-1, 6)
    h = len(coords) // 2
    for i, v in enumerate(coords[:h]):
        x[i] = L.fp_from_int(v)
    for i, v in enumerate(coords[h:]):
        y[i] = L.fp_from_int(v)
    return p


# generator coordinates (g1.go:25-26, g2.go:26-29): plain data, converted to Montgomery limbs by layout.fp_from_int
_G1_GEN = _gen_pod(L.G1_AFFINE, [
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1])
_G2_GEN = _gen_pod(L.G2_AFFINE, [
    0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
    0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e,
    0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
    0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be])


def product_is_one(p_list, q_list):
    """FinalExponentiation(prod MillerLoop(P_i, Q_i)) == 1 on the engine (CompareTwoPairings generalised, pairing.go:140-147)"""
    P = np.concatenate(p_list); Q = np.concatenate(q_list)
    return bool(engine().pairing_product_is_one(P, Q, [0, P.size])[0])


class SecretKey:
    def __init__(self, f):
        self.f = f % L.R_ORDER                                   # FR element (canonical integer)

    def Serialize(self):                                         # g1pubs/bls.go:115-118
        return self.f.to_bytes(32, "big")


def DeserializeSecretKey(b):                                     # g1pubs/bls.go:121-123 (FRReprToFR returns nil for values >= r)
    v = int.from_bytes(bytes(b), "big")
    return SecretKey(v) if v < L.R_ORDER else None


def DeriveSecretKey(b32):                                        # g1pubs/bls.go:127-129
    return SecretKey(K.hash_secret_key(b32))


def RandKey(reader):                                             # g1pubs/bls.go:149-156
    return SecretKey(K.rand_int(reader, L.R_ORDER))


def build(ns, key_group, doc_prefix):
    """populate module namespace `ns` with the g1pubs (key_group = 1) or g2pubs (key_group = 2) API"""
    KG, SG = Group(key_group), Group(3 - key_group)

    def pairs(key_pt, sig_pt):
        """order a (key-group point, signature-group point) pair as (G1, G2) for the Miller loop"""
        return (key_pt, sig_pt) if key_group == 1 else (sig_pt, key_pt)

    class PublicKey:
        def __init__(self, p):
            self.p = p

        def Serialize(self):
            return KG.compress(self.p)

        def Equals(self, other):
            return self.p.tobytes() == other.p.tobytes()

        def Copy(self):
            return PublicKey(self.p.copy())

        def Aggregate(self, other):
            self.p = KG.sum([self.p, other.p])

    class Signature:
        def __init__(self, s):
            self.s = s

        def Serialize(self):
            return SG.compress(self.s)

        def Copy(self):
            return Signature(self.s.copy())

        def Aggregate(self, other):
            self.s = SG.sum([self.s, other.s])

        # e(G_key, sig) == prod e(pk_i, H(m_i))  <=>  FE(ML(-G_key, sig) * prod ML(pk_i, H(m_i))) == 1
        def VerifyAggregate(self, pubKeys, msgs):                # g1pubs/bls.go:252-282, g2pubs/bls.go:240-270
            if len(pubKeys) != len(msgs):
                return False
            last = b""                                           # Go: bytes.Equal(m, nil) is true for an empty message
            for m in sorted(bytes(m) for m in msgs):
                if m == last:
                    return False
                last = m
            if _is_inf(self.s) or any(_is_inf(pk.p) for pk in pubKeys):
                return False                                     # fail closed: the reference panics on these (pairing.go:17-26)
            hs = (engine().hash_g2_batch if key_group == 1 else engine().hash_g1_batch)([bytes(m) for m in msgs])
            prs = [pairs(KG.neg(KG.generator()), self.s)] + [pairs(pk.p, hs[i:i + 1]) for i, pk in enumerate(pubKeys)]
            return product_is_one([a for a, _ in prs], [b for _, b in prs])

        def VerifyAggregateCommon(self, pubKeys, msg):           # g1pubs/bls.go:287-290
            return Verify(msg, AggregatePublicKeys(pubKeys), self)

    def _is_inf(pt):
        return bool(np.asarray(pt["inf"]).any())

    def _check(pub, sig, h):
        """CompareTwoPairings(G_key, sig, pub, h) (pairing.go:140-147).  An infinite key or signature is rejected: the engine's
        Miller loop treats a pair with a point at infinity as the factor 1 (the reference panics there, pairing.go:17-26), so
        without this guard pub = sig = infinity would verify for every message."""
        if _is_inf(pub.p) or _is_inf(sig.s):
            return False
        prs = [pairs(KG.generator(), sig.s), pairs(KG.neg(pub.p), h)]
        return product_is_one([a for a, _ in prs], [b for _, b in prs])

    def DeserializeSignature(b):
        return Signature(SG.decompress(b))

    def DeserializePublicKey(b):
        return PublicKey(KG.decompress(b))

    def PrivToPub(k):
        return PublicKey(KG.mul(KG.generator(), k.f))

    def Sign(message, key):
        return Signature(SG.mul(SG.hash(message), key.f))

    def Verify(m, pub, sig):
        return _check(pub, sig, SG.hash(m))

    def AggregateSignatures(sigs):
        return Signature(SG.sum([s.s for s in sigs]))

    def AggregatePublicKeys(pubs):
        return PublicKey(KG.sum([p.p for p in pubs]))

    def NewAggregateSignature():
        return Signature(SG.zero())

    def NewAggregatePubkey():
        return PublicKey(KG.zero())

    ns.update(PublicKey=PublicKey, Signature=Signature, DeserializeSignature=DeserializeSignature,
              DeserializePublicKey=DeserializePublicKey, PrivToPub=PrivToPub, Sign=Sign, Verify=Verify,
              AggregateSignatures=AggregateSignatures, AggregatePublicKeys=AggregatePublicKeys,
              NewAggregateSignature=NewAggregateSignature, NewAggregatePubkey=NewAggregatePubkey,
              SecretKey=SecretKey, DeserializeSecretKey=DeserializeSecretKey, DeriveSecretKey=DeriveSecretKey, RandKey=RandKey,
              engine=engine, set_engine=set_engine)
    if key_group == 1:            # the WithDomain variants exist only in g1pubs (g1pubs/bls.go:138-141,171-174,294-311)
        def _hash_domain(m32, domain8):
            return engine().hash_g2_with_domain_batch([bytes(m32)], bytes(domain8))

        def SignWithDomain(message32, key, domain8):
            return Signature(SG.mul(_hash_domain(message32, domain8), key.f))

        def VerifyWithDomain(m32, pub, sig, domain8):
            return _check(pub, sig, _hash_domain(m32, domain8))

        def VerifyAggregateCommonWithDomain(self, pubKeys, msg32, domain8):
            return VerifyWithDomain(msg32, AggregatePublicKeys(pubKeys), self, domain8)

        def VerifyAggregateWithDomain(self, pubKeys, msgs32, domain8):
            if len(pubKeys) != len(msgs32):
                return False
            if _is_inf(self.s) or any(_is_inf(pk.p) for pk in pubKeys):
                return False
            hs = engine().hash_g2_with_domain_batch([bytes(m) for m in msgs32], bytes(domain8))
            return product_is_one([KG.neg(KG.generator())] + [pk.p for pk in pubKeys], [self.s] + [hs[i:i + 1] for i in range(len(pubKeys))])

        Signature.VerifyAggregateCommonWithDomain = VerifyAggregateCommonWithDomain
        Signature.VerifyAggregateWithDomain = VerifyAggregateWithDomain
        ns.update(SignWithDomain=SignWithDomain, VerifyWithDomain=VerifyWithDomain)

        # batch additions (the reference has none): n wire-format triples per call, everything on the device
        def VerifyBatch(pubs48, msgs, sigs96):
            return engine().g1pubs_verify_batch(np.frombuffer(b"".join(pubs48), np.uint8), msgs, np.frombuffer(b"".join(sigs96), np.uint8)).astype(bool).tolist()

        def VerifyWithDomainBatch(pubs48, msgs32, domain8, sigs96):
            return engine().verify_with_domain_batch(pubs48, msgs32, domain8, sigs96).astype(bool).tolist()
        ns.update(VerifyBatch=VerifyBatch, VerifyWithDomainBatch=VerifyWithDomainBatch)
    else:
        def VerifyBatch(pubs96, msgs, sigs48):
            return engine().g2pubs_verify_batch(np.frombuffer(b"".join(pubs96), np.uint8), msgs, np.frombuffer(b"".join(sigs48), np.uint8)).astype(bool).tolist()
        ns.update(VerifyBatch=VerifyBatch)
