"""Python mirror of package g2pubs (public keys in G2, signatures in G1; g2pubs/bls.go) on top of the engine.
See g1pubs.py for the conventions; BASELINE.json's first configuration (g2pubs Sign / Verify) is this module."""
from . import hostgen as hg, hostmath as hm, layout as L
from .g1pubs import engine, set_engine, SecretKey, DeriveSecretKey, DeserializeSecretKey, RandKey, _g1_from_jac, _g2_from_jac  # noqa: F401


class PublicKey:
    def __init__(self, p):
        self.p = p                                               # G2 point

    def Serialize(self):                                         # g2pubs/bls.go:67-69
        return hm.compress_g2(self.p)

    def Equals(self, other):                                     # :85-88
        return self.Serialize() == other.Serialize()

    def Copy(self):
        return PublicKey(self.p)

    def Aggregate(self, other):                                  # :189-192
        self.p = hm.g2_add(self.p, other.p)


class Signature:
    def __init__(self, s):
        self.s = s                                               # G1 point

    def Serialize(self):                                         # :18-20
        return hm.compress_g1(self.s)

    def Copy(self):
        return Signature(self.s)

    def Aggregate(self, other):                                  # :174-177
        self.s = hm.g1_add(self.s, other.s)

    def VerifyAggregate(self, pubKeys, msgs):                    # :240-270
        if len(pubKeys) != len(msgs):
            return False
        last = b""
        for m in sorted(bytes(m) for m in msgs):
            if m == last:
                return False
            last = m
        # e(sig, G2One) == prod e(H(m_i), pk_i)
        return _product_is_one([(hm.g1_neg(self.s), hm.G2)] + [(hm.hash_g1(m), pk.p) for pk, m in zip(pubKeys, msgs)])

    def VerifyAggregateCommon(self, pubKeys, msg):               # :275-278
        return Verify(msg, AggregatePublicKeys(pubKeys), self)


def _product_is_one(pairs):
    P = hg.g1_points([p for p, _ in pairs]); Q = hg.g2_points([q for _, q in pairs])
    return bool(engine().pairing_product_is_one(P, Q, [0, len(pairs)])[0])


def DeserializeSignature(b):                                     # :33-40
    p, err = hm.decompress_g1(b)
    if err:
        raise ValueError(err)
    return Signature(p)


def DeserializePublicKey(b):                                     # :91-98
    p, err = hm.decompress_g2(b)
    if err:
        raise ValueError(err)
    return PublicKey(p)


def PrivToPub(k):                                                # :138-140
    return PublicKey(hm.g2_mul(hm.G2, k.f))


def Sign(message, key):                                          # :132-135
    return Signature(hm.g1_mul(hm.hash_g1(message), key.f))


def Verify(m, pub, sig):                                         # :159-162: CompareTwoPairings(sig, G2One, HashG1(m), pub)
    return _product_is_one([(sig.s, hm.G2), (hm.g1_neg(hm.hash_g1(m)), pub.p)])


def AggregateSignatures(sigs):                                   # :165-171 -> b381_g1_sum
    return Signature(_g1_from_jac(engine().g1_sum(hg.g1_points([s.s for s in sigs]))))


def AggregatePublicKeys(pubs):                                   # :180-186 -> b381_g2_sum
    return PublicKey(_g2_from_jac(engine().g2_sum(hg.g2_points([p.p for p in pubs]))))


def NewAggregateSignature():
    return Signature(None)


def NewAggregatePubkey():
    return PublicKey(None)
