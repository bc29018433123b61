"""Shared inputs for the wire-format tests (tests/test_emu_codec.py on the CPU emulation of csrc/codec.cuh,
tests/test_gpu_codec.py on the GPU): valid points, the reference's invalid vectors, and hand-made edge cases."""
import numpy as np

from bls_b200 import layout as L

# TestPubkeyDeserializeInvalid: g1pubs/bls_test.go:422-433 and g2pubs/bls_test.go:336-347
REF_INVALID_G1 = bytes.fromhex("b5a44e98d450f266567be0d82e60d965aa8703f73a9a71aa03b98215444f781d00000000000000000000000000000000")
REF_INVALID_G2 = bytes.fromhex("b5a44e98e450f266567be0d82e60d965aa8703f73a9a71aa03b98215444f781d00000000000000000000000000000000"
                               "b5a44e98d450f266567be0d82e60d965aa8703f73a9a71aa03b98215444f781d00000000000000000000000000000000")


def compressed_cases(orc, grp, nbytes, points, seed):
    """list of compressed encodings: valid points (both y signs), infinity, and malformed / off-curve / off-subgroup ones"""
    rng = np.random.RandomState(seed)
    good = [grp.compress(points[i:i + 1]) for i in range(points.size)]
    cases = list(good)
    inf = bytearray(nbytes); inf[0] = 0xc0
    cases.append(bytes(inf))
    bad = bytearray(inf); bad[-1] = 1; cases.append(bytes(bad))                 # junk in a compressed infinity
    bad = bytearray(inf); bad[0] = 0xe0; cases.append(bytes(bad))               # infinity with the sign bit
    bad = bytearray(good[0]); bad[0] &= 0x7f; cases.append(bytes(bad))          # compression bit missing
    flip = bytearray(good[1]); flip[0] ^= 0x20; cases.append(bytes(flip))       # the other root
    # x >= Q decodes as 0 (FQReprToFQ, fq.go:49-56): all-ones coordinate
    big = bytearray(b"\xff" * nbytes); big[0] = 0x9f; cases.append(bytes(big))
    big = bytearray(b"\xff" * nbytes); big[0] = 0xbf; cases.append(bytes(big))
    # x = Q exactly, x = Q - 1, x = 0, x = 1 ...
    for v in (L.Q, L.Q - 1, 0, 1, 2, 3, 4, 5):
        b = bytearray(v.to_bytes(48, "big").rjust(nbytes, b"\0")) if nbytes == 48 else bytearray(v.to_bytes(48, "big") + (v // 3).to_bytes(48, "big"))
        b[0] |= 0x80; cases.append(bytes(b))
        b[0] |= 0x20; cases.append(bytes(b))
    # random x: about half are not on the curve, the rest are on the curve but (almost surely) outside the r-torsion
    for _ in range(24):
        b = bytearray(rng.randint(0, 256, nbytes, dtype=np.uint8).tobytes())
        b[0] = (b[0] & 0x1f) | 0x80 | (0x20 if rng.randint(2) else 0)
        if nbytes == 96:
            b[48] &= 0x1f
        cases.append(bytes(b))
    return cases


def oracle_decompress(grp, dtype, cases, checked):
    pts = np.zeros(len(cases), dtype=dtype); st = np.zeros(len(cases), np.uint8)
    for i, c in enumerate(cases):
        e, o = grp.decompress(c, checked)
        st[i] = e
        if e in (0, 4) or True:
            pts[i] = o[0]
    return pts, st
