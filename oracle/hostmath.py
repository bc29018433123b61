"""TEST INFRASTRUCTURE (oracle): pure-Python-integer restatement of the reference's hashing to the curve (hash.go,
g1.go:614-714, g2.go:883-1085), point (de)compression, key derivation and the group law, used ONLY by tests/ to
check the GPU kernels of csrc/hash.cuh, csrc/swu.cuh and csrc/codec.cuh.  Nothing under bls_b200/ imports it.
Pinned by the reference's own known-answer tests (tests/test_host_api.py: hash_test.go:12-82,
g1pubs/bls_test.go:409-420) and cross-checked against the C++ oracle (compression, subgroup membership).

Values are canonical integers (NOT Montgomery); Fq2 elements are (c0, c1) tuples; points are affine
tuples (x, y) or None for infinity.  Conversion to the engine's Montgomery PODs is in hostgen.g1_points /
g2_points.
"""
import hashlib
import json
import os

from bls_b200 import layout as L
from bls_b200.hostgen import _Fq, _Fq2, _dbl, _madd, _mul, _to_affine_batch, G1, G2

Q = L.Q
R_ORDER = L.R_ORDER
BLS_X = L.BLS_X

with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bls_b200", "hash_params.json")) as _f:
    _P = json.load(_f)
_h = lambda s: int(s, 16)
ISO11 = [[_h(c) for c in _P["iso11"][n]] for n in ("xNum11", "xDen11", "yNum11", "yDen11")]
ISO3 = [[(_h(a), _h(b)) for a, b in _P["iso3"][n]] for n in ("xNum3", "xDen3", "yNum3", "yDen3")]
IWSC = (_h(_P["iwsc"][0]), _h(_P["iwsc"][1]))
KQIX, KQIY = _h(_P["kQiX"]), _h(_P["kQiY"])
ELLPA, ELLPB = _h(_P["ellPA"]), _h(_P["ellPB"])
ELL2PA, ELL2PB = tuple(_P["ell2pA"]), tuple(_P["ell2pB"])
G2_COFACTOR = _h(_P["g2_cofactor"])
NQR2 = (1, 1)                      # fq2nqr = 1 + u (fq6.go:139-142)
Q_MINUS_1_OVER_2 = (Q - 1) // 2


# ---- field helpers ----------------------------------------------------------------------------------------
def fq_sqrt(a):
    """FQ.Sqrt (fq.go:203-217): a^((q+1)/4) when a is a square"""
    a1 = pow(a, (Q - 3) // 4, Q)
    if a1 * a1 * a % Q == Q - 1:
        return None
    return a1 * a % Q


def fq2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _Fq2.mul(r, a)
        a = _Fq2.sqr(a)
        e >>= 1
    return r


def fq2_sqrt(a):
    """FQ2.Sqrt (fq2.go:198-232): algorithm 9 of eprint 2012/685"""
    if a == (0, 0):
        return (0, 0)
    a1 = fq2_pow(a, (Q - 3) // 4)
    alpha = _Fq2.mul(_Fq2.sqr(a1), a)
    a0 = _Fq2.mul((alpha[0], (-alpha[1]) % Q), alpha)
    neg1 = (Q - 1, 0)
    if a0 == neg1:
        return None
    a1 = _Fq2.mul(a1, a)
    if alpha == neg1:
        return _Fq2.mul(a1, (0, 1))
    alpha = fq2_pow(_Fq2.add(alpha, (1, 0)), Q_MINUS_1_OVER_2)
    return _Fq2.mul(alpha, a1)


def fq2_cmp(a, b):
    """FQ2.Cmp (fq2.go:31-37): by c1, then c0"""
    return (a[1] > b[1]) - (a[1] < b[1]) or (a[0] > b[0]) - (a[0] < b[0])


def fq2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


# ---- groups -------------------------------------------------------------------------------------------------
def _affine(F, P):
    return _to_affine_batch(F, [P])[0]


def g1_mul(p, k):
    """affine p times integer k (MSB-first double-and-add as G1Affine.Mul / MulFR, g1.go:67-90)"""
    return None if p is None else _affine(_Fq, _mul(_Fq, p, k))


def g2_mul(p, k):
    return None if p is None else _affine(_Fq2, _mul(_Fq2, p, k))


def _add_affine(F, a, b):
    if a is None:
        return b
    if b is None:
        return a
    return _affine(F, _madd(F, (a[0], a[1], F.one), b))


def g1_add(a, b): return _add_affine(_Fq, a, b)
def g2_add(a, b): return _add_affine(_Fq2, a, b)
def g1_neg(p): return None if p is None else (p[0], (-p[1]) % Q)
def g2_neg(p): return None if p is None else (p[0], fq2_neg(p[1]))


def g1_in_subgroup(p):
    return g1_mul(p, R_ORDER) is None        # g1.go:137-141


def g2_in_subgroup(p):
    return g2_mul(p, R_ORDER) is None        # g2.go:293-295


# ---- compression (g1.go:185-249, g2.go:219-289) ----------------------------------------------------------
def compress_g1(p):
    if p is None:
        return bytes([0xC0]) + bytes(47)
    out = bytearray(p[0].to_bytes(48, "big"))
    if p[1] > (-p[1]) % Q:
        out[0] |= 1 << 5
    out[0] |= 1 << 7
    return bytes(out)


def compress_g2(p):
    if p is None:
        return bytes([0xC0]) + bytes(95)
    out = bytearray(p[0][1].to_bytes(48, "big") + p[0][0].to_bytes(48, "big"))     # x.c1 first
    if fq2_cmp(p[1], fq2_neg(p[1])) > 0:
        out[0] |= 1 << 5
    out[0] |= 1 << 7
    return bytes(out)


def decompress_g1(b, checked=True):
    """-> (point, error string or None), the error texts of g1.go:185-227"""
    b = bytearray(b)
    if not b[0] & 0x80:
        return None, "unexpected compression mode"
    if b[0] & 0x40:
        b[0] &= 0x3F
        if any(b):
            return None, "unexpected information in compressed infinity"
        return None, None
    greatest = bool(b[0] & 0x20)
    b[0] &= 0x1F
    x = int.from_bytes(b, "big")
    if x >= Q:
        return None, "not in field"
    y = fq_sqrt((x * x * x + 4) % Q)
    if y is None:
        return None, "point not on curve"
    ny = (-y) % Q
    p = (x, y if (y < ny) != greatest else ny)
    if checked and not g1_in_subgroup(p):
        return None, "not in correct subgroup"
    return p, None


def decompress_g2(b, checked=True):
    b = bytearray(b)
    if not b[0] & 0x80:
        return None, "unexpected compression mode"
    if b[0] & 0x40:
        b[0] &= 0x3F
        if any(b):
            return None, "unexpected information in infinity point on G2"
        return None, None
    greatest = bool(b[0] & 0x20)
    b[0] &= 0x1F
    x = (int.from_bytes(b[48:], "big"), int.from_bytes(b[:48], "big"))
    if x[0] >= Q or x[1] >= Q:
        return None, "not in field"
    y = fq2_sqrt(_Fq2.add(_Fq2.mul(_Fq2.sqr(x), x), (4, 4)))
    if y is None:
        return None, "point not on curve"
    ny = fq2_neg(y)
    p = (x, y if (fq2_cmp(y, ny) < 0) != greatest else ny)
    if checked and not g2_in_subgroup(p):
        return None, "point is not in correct subgroup"
    return p, None


# ---- hash to field (hash.go:9-113) ------------------------------------------------------------------------
def _sha(*parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(p)
    return h.digest()


def hash_secret_key(b32):
    """HashSecretKey (hash.go:9-39) -> integer mod r"""
    prime = _sha(bytes(b32)) + b"\x00"
    t = b"".join(_sha(prime, b"\x01", bytes([j])) for j in (1, 2))
    return int.from_bytes(t, "big") % R_ORDER


def _hp(msg, ctr):
    prime = _sha(msg) + bytes([ctr])
    t = b"".join(_sha(prime, b"\x01", bytes([j])) for j in (1, 2))
    return int.from_bytes(t, "big") % Q


def _hp2(msg, ctr):
    prime = _sha(msg) + bytes([ctr])
    return tuple(int.from_bytes(b"".join(_sha(prime, bytes([i]), bytes([j])) for j in (1, 2)), "big") % Q for i in (1, 2))


# ---- simplified SWU + isogenies + cofactor clearing (g1.go:614-714, g2.go:883-1031, hash.go:115-411) ----------
def _sign_fq(f):
    return -1 if f > Q_MINUS_1_OVER_2 else 1


def _swu_g1(t):
    inv = lambda v: pow(v, -1, Q)
    t2 = t * t % Q
    common = (t2 * t2 - t2) % Q                       # xi^2 t^4 + xi t^2 with xi = -1
    if common == 0:
        x0 = ELLPB * inv((-ELLPA) % Q) % Q
    else:
        x0 = (-ELLPB) % Q * ((common + 1) % Q) % Q * inv(ELLPA * common % Q) % Q
    g = lambda x: (x * x * x + ELLPA * x + ELLPB) % Q
    y = fq_sqrt(g(x0))
    x = x0
    if y is None:
        x = (-t2) % Q * x0 % Q
        y = fq_sqrt(g(x))
        assert y is not None
    if _sign_fq(y) * _sign_fq(t) < 0:
        y = (-y) % Q
    return (x, y)


def _horner(F, coeffs, x):
    r = coeffs[-1]
    for c in reversed(coeffs[:-1]):
        r = F.add(F.mul(r, x), c)
    return r


def _iso(F, maps, p):
    x, y = p
    xn, xd, yn, yd = (_horner(F, m, x) for m in maps)
    return (F.mul(xn, F.inv(xd)), F.mul(F.mul(y, yn), F.inv(yd)))


def hash_g1(msg):
    """bls.HashG1 (hash.go:320-331)"""
    m = b"\x01" + bytes(msg)
    # the reference adds the two SWU points with its ordinary AddAffine (hash.go:312-318 -> g1.go:485-559)
    p = _add_affine(_Fq, _swu_g1(_hp(m, 0)), _swu_g1(_hp(m, 1)))
    p = _iso(_Fq, ISO11, p)
    return g1_add(g1_mul(p, BLS_X), p)                # ClearH (hash.go:305-309)


def _small(F, k):
    return k if F is _Fq else (k, 0)


def _sign_fq2(f):
    """signFQ2 (g2.go:920-934)"""
    if f[1] > Q_MINUS_1_OVER_2:
        return -1
    if f[1] > 0:
        return 1
    if f[0] > Q_MINUS_1_OVER_2:
        return -1
    return 1


def _swu_g2(t):
    F = _Fq2
    t2 = F.sqr(t)
    common = F.add(F.mul(F.sqr(NQR2), F.sqr(t2)), F.mul(NQR2, t2))
    if common == (0, 0):
        x0 = F.mul(ELL2PB, F.inv(F.mul(NQR2, ELL2PA)))
    else:
        x0 = F.mul(F.mul(fq2_neg(ELL2PB), F.add(common, (1, 0))), F.inv(F.mul(ELL2PA, common)))
    gx0 = F.add(F.add(F.mul(F.sqr(x0), x0), F.mul(ELL2PA, x0)), ELL2PB)
    y = fq2_sqrt(gx0)
    if y is not None and F.sqr(y) == gx0:
        if _sign_fq2(t) != _sign_fq2(y):
            y = fq2_neg(y)
        return (x0, y)
    t6 = F.sqr(F.mul(t2, t))
    x1 = F.mul(F.mul(NQR2, t2), x0)
    gx1 = F.mul(F.mul(F.mul(F.sqr(NQR2), NQR2), t6), gx0)
    y = fq2_sqrt(gx1)
    assert y is not None and F.sqr(y) == gx1
    if _sign_fq2(t) != _sign_fq2(y):
        y = fq2_neg(y)
    return (x1, y)


def _psi(p):
    """hash.go:341-366"""
    F = _Fq2
    qx = F.mul(IWSC, p[0])
    qx = (qx[0] * KQIX % Q, (-(qx[1] * KQIX)) % Q)
    nx = F.mul(NQR2, qx)
    qy = F.mul(IWSC, p[1])
    qy = ((qy[0] + qy[1]) * KQIY % Q, (qy[0] - qy[1]) * KQIY % Q)
    return (nx, F.mul(NQR2, qy))


def _clear_h2(p):
    """hash.go:368-389"""
    work = g2_add(g2_mul(p, BLS_X), p)
    mpsi = g2_neg(_psi(p))
    work = g2_add(work, mpsi)
    work = g2_add(g2_mul(work, BLS_X), mpsi)
    work = g2_add(work, g2_neg(p))
    p2 = _affine(_Fq2, _dbl(_Fq2, (p[0], p[1], _Fq2.one)))
    return g2_add(work, _psi(_psi(p2)))


def hash_g2(msg):
    """bls.HashG2 (hash.go:404-411)"""
    m = b"\x01" + bytes(msg)
    p = _add_affine(_Fq2, _swu_g2(_hp2(m, 0)), _swu_g2(_hp2(m, 1)))
    return _clear_h2(_iso(_Fq2, ISO3, p))


def hash_g2_with_domain(message_hash32, domain8):
    """bls.HashG2WithDomain (g2.go:1041-1085): try-and-increment, lower y, times the G2 cofactor"""
    base = bytes(message_hash32) + bytes(domain8)
    x = (int.from_bytes(_sha(base + b"\x01"), "big"), int.from_bytes(_sha(base + b"\x02"), "big"))
    while True:
        y = fq2_sqrt(_Fq2.add(_Fq2.mul(_Fq2.sqr(x), x), (4, 4)))
        if y is not None:
            if not fq2_cmp(y, fq2_neg(y)) > 0:          # Parity(): keep the value that is larger than its negative
                y = fq2_neg(y)
            return g2_mul((x, y), G2_COFACTOR)
        x = _Fq2.add(x, (1, 0))


# ---- the tests' deterministic key source (g1_test.go:106-124 + Go's crypto/rand.Int) ------------------------------
class XorShiftReader:
    def __init__(self, seed):
        self.x = seed & 0xFFFFFFFFFFFFFFFF

    def read(self, n):
        out = bytearray()
        for _ in range(n):
            x = self.x
            x ^= (x << 13) & 0xFFFFFFFFFFFFFFFF
            x ^= x >> 7
            x ^= (x << 17) & 0xFFFFFFFFFFFFFFFF
            self.x = x
            out.append(x & 0xFF)
        return bytes(out)


def rand_int(reader, maximum):
    """crypto/rand.Int(reader, max): rejection sampling of ceil(bitlen/8) bytes with the top byte masked"""
    n = maximum - 1
    bl = n.bit_length()
    k = (bl + 7) // 8
    b = bl % 8 or 8
    while True:
        raw = bytearray(reader.read(k))
        raw[0] &= (1 << b) - 1
        v = int.from_bytes(raw, "big")
        if v < maximum:
            return v
