"""Flat POD layouts shared by the C ABI (include/b381.h) and its Python callers.

All field elements are 6 x u64 limbs, least-significant limb first, Montgomery form with
R = 2^384, canonical in [0, Q) -- the reference's in-memory format (fqrepr.go:13-14, fq.go:41-45).
"""
import numpy as np

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab  # fq.go:26
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001  # fr.go:16
MONT_R = (1 << 384) % Q
MONT_R_INV = pow(MONT_R, -1, Q)
BLS_X = 0xd201000000010000  # g2.go:634 (the parameter is -BLS_X, g2.go:636)

U64 = np.uint64
FP_BYTES, FP2_BYTES, FP12_BYTES = 48, 96, 576

#: G1Affine{x, y FQ; infinity bool} with Go's padding (g1.go:10-14): 104 bytes
G1_AFFINE = np.dtype([("x", U64, (6,)), ("y", U64, (6,)), ("inf", np.uint8), ("pad", np.uint8, (7,))])
#: G2Affine{x, y FQ2; infinity bool} (g2.go:12-16): 200 bytes
G2_AFFINE = np.dtype([("x", U64, (2, 6)), ("y", U64, (2, 6)), ("inf", np.uint8), ("pad", np.uint8, (7,))])
#: G1Projective{x, y, z FQ} (g1.go:252-256): 144 bytes
G1_JAC = np.dtype([("x", U64, (6,)), ("y", U64, (6,)), ("z", U64, (6,))])
#: G2Projective{x, y, z FQ2} (g2.go:298-302): 288 bytes
G2_JAC = np.dtype([("x", U64, (2, 6)), ("y", U64, (2, 6)), ("z", U64, (2, 6))])
#: FQ12 flattened c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2, each FQ2 = c0 || c1 (fq12.go:9-12): 576 bytes
FP12 = np.dtype((U64, (2, 3, 2, 6)))
# G2Prepared (g2.go:639-642): 68 coefficient triples (c0 || c1 each), infinity flag + padding = b381_g2_prepared
G2_PREPARED = np.dtype([("coeffs", U64, (68, 3, 2, 6)), ("inf", np.uint8), ("pad", np.uint8, (7,))])
#: canonical scalar < r, 4 x u64 LS limb first (frrepr.go:11)
SCALAR = np.dtype((U64, (4,)))

assert G1_AFFINE.itemsize == 104 and G2_AFFINE.itemsize == 200
assert G1_JAC.itemsize == 144 and G2_JAC.itemsize == 288 and FP12.itemsize == 576


def int_to_limbs(v, n=6):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def limbs_to_int(l):
    return sum(int(x) << (64 * i) for i, x in enumerate(l))


def fp_from_int(v):
    """canonical integer -> Montgomery limbs (what FQReprToFQ does, fq.go:49-56)"""
    return np.array(int_to_limbs(v % Q * MONT_R % Q), dtype=U64)


def fp_to_int(limbs):
    """Montgomery limbs -> canonical integer (FQ.ToRepr, fq.go:334-338)"""
    return limbs_to_int(limbs) * MONT_R_INV % Q


def scalar_from_int(v):
    return np.array(int_to_limbs(v, 4), dtype=U64)


def scalar_to_int(l):
    return limbs_to_int(l)
