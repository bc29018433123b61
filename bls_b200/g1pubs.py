"""Python mirror of package g1pubs (public keys in G1, signatures in G2; g1pubs/bls.go) on top of the engine:
same names, argument meaning and results as the reference, with the pairing and aggregation work done by
libb381.so through the C ABI and the host-side work (hashing, serialisation, scalar multiplication for
signing) in hostmath.py -- the split the Go shim of INTEGRATION.md makes.

Points are kept as affine integer tuples (hostmath conventions); `engine()` is the process-wide b381 context.
"""
from . import hostgen as hg, hostmath as hm, layout as L

_ctx = None


def engine():
    global _ctx
    if _ctx is None:
        from . import capi
        _ctx = capi.Ctx(0)
    return _ctx


def set_engine(ctx):
    global _ctx
    _ctx = ctx


class SecretKey:
    def __init__(self, f):
        self.f = f % L.R_ORDER                                   # FR element (canonical integer)

    def Serialize(self):                                         # g1pubs/bls.go:115-118
        return self.f.to_bytes(32, "big")


class PublicKey:
    def __init__(self, p):
        self.p = p                                               # G1 point

    def Serialize(self):                                         # :67-69
        return hm.compress_g1(self.p)

    def Equals(self, other):
        return self.p == other.p

    def Copy(self):
        return PublicKey(self.p)

    def Aggregate(self, other):                                  # :201-204
        self.p = hm.g1_add(self.p, other.p)


class Signature:
    def __init__(self, s):
        self.s = s                                               # G2 point

    def Serialize(self):                                         # :18-20
        return hm.compress_g2(self.s)

    def Copy(self):
        return Signature(self.s)

    def Aggregate(self, other):                                  # :186-189
        self.s = hm.g2_add(self.s, other.s)

    # ---- verification (the hot path: on the GPU) ---------------------------------------------------------
    def VerifyAggregate(self, pubKeys, msgs):                    # :252-282
        if len(pubKeys) != len(msgs):
            return False
        last = b""                                               # Go: bytes.Equal(m, nil) is true for an empty message
        for m in sorted(bytes(m) for m in msgs):
            if m == last:
                return False
            last = m
        return _product_is_one([(hm.g1_neg(hm.G1), self.s)] + [(pk.p, hm.hash_g2(m)) for pk, m in zip(pubKeys, msgs)])

    def VerifyAggregateCommon(self, pubKeys, msg):               # :287-290
        return Verify(msg, AggregatePublicKeys(pubKeys), self)

    def VerifyAggregateCommonWithDomain(self, pubKeys, msg32, domain8):   # :294-297
        return VerifyWithDomain(msg32, AggregatePublicKeys(pubKeys), self, domain8)

    def VerifyAggregateWithDomain(self, pubKeys, msgs32, domain8):        # :300-311
        if len(pubKeys) != len(msgs32):
            return False
        return _product_is_one([(hm.g1_neg(hm.G1), self.s)] +
                               [(pk.p, hm.hash_g2_with_domain(m, domain8)) for pk, m in zip(pubKeys, msgs32)])


def _product_is_one(pairs):
    """FinalExponentiation(prod MillerLoop(P_i, Q_i)) == 1 on the engine; an infinity pair contributes 1"""
    P = hg.g1_points([p for p, _ in pairs]); Q = hg.g2_points([q for _, q in pairs])
    return bool(engine().pairing_product_is_one(P, Q, [0, len(pairs)])[0])


def DeserializeSignature(b):                                     # :38-45
    p, err = hm.decompress_g2(b)
    if err:
        raise ValueError(err)
    return Signature(p)


def DeserializePublicKey(b):                                     # :91-98
    p, err = hm.decompress_g1(b)
    if err:
        raise ValueError(err)
    return PublicKey(p)


def DeserializeSecretKey(b):                                     # :121-123 (FRReprToFR returns nil for values >= r)
    v = int.from_bytes(bytes(b), "big")
    return SecretKey(v) if v < L.R_ORDER else None


def DeriveSecretKey(b32):                                        # :127-129
    return SecretKey(hm.hash_secret_key(b32))


def RandKey(reader):                                             # :149-156
    return SecretKey(hm.rand_int(reader, L.R_ORDER))


def PrivToPub(k):                                                # :144-146
    return PublicKey(hm.g1_mul(hm.G1, k.f))


def Sign(message, key):                                          # :132-135
    return Signature(hm.g2_mul(hm.hash_g2(message), key.f))


def SignWithDomain(message32, key, domain8):                     # :138-141
    return Signature(hm.g2_mul(hm.hash_g2_with_domain(message32, domain8), key.f))


def Verify(m, pub, sig):                                         # :165-168: CompareTwoPairings(G1One, sig, pub, H(m))
    return _product_is_one([(hm.G1, sig.s), (hm.g1_neg(pub.p), hm.hash_g2(m))])


def VerifyWithDomain(m32, pub, sig, domain8):                    # :171-174
    return _product_is_one([(hm.G1, sig.s), (hm.g1_neg(pub.p), hm.hash_g2_with_domain(m32, domain8))])


def AggregateSignatures(sigs):                                   # :177-183 -> b381_g2_sum
    out = engine().g2_sum(hg.g2_points([s.s for s in sigs]))
    return Signature(_g2_from_jac(out))


def AggregatePublicKeys(pubs):                                   # :192-198 -> b381_g1_sum
    out = engine().g1_sum(hg.g1_points([p.p for p in pubs]))
    return PublicKey(_g1_from_jac(out))


def NewAggregateSignature():
    return Signature(None)


def NewAggregatePubkey():
    return PublicKey(None)


def _g1_from_jac(j):
    if not j["z"].any():
        return None
    return (L.fp_to_int(j["x"][0]), L.fp_to_int(j["y"][0]))      # the engine returns z = 1


def _g2_from_jac(j):
    if not j["z"].any():
        return None
    return ((L.fp_to_int(j["x"][0][0]), L.fp_to_int(j["x"][0][1])), (L.fp_to_int(j["y"][0][0]), L.fp_to_int(j["y"][0][1])))
