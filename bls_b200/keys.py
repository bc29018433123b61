"""Secret-key handling of g1pubs / g2pubs that involves no curve arithmetic (the Go host keeps doing it with the
standard library): HashSecretKey (hash.go:9-39), crypto/rand.Int as RandKey uses it (g1pubs/bls.go:149-156,
fr.go:337-344), and the xorshift byte source of the reference's tests (g1_test.go:106-124)."""
import hashlib

from . import layout as L


def hash_secret_key(b32):
    """HashSecretKey (hash.go:9-39): hash_to_field with m = 1 over the scalar field -> integer mod r"""
    prime = hashlib.sha256(bytes(b32)).digest() + b"\x00"
    t = b"".join(hashlib.sha256(prime + b"\x01" + bytes([j])).digest() for j in (1, 2))
    return int.from_bytes(t, "big") % L.R_ORDER


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


class XorShiftReader:
    """NewXORShift(seed) (g1_test.go:106-124): the deterministic byte source of the reference's tests"""

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
