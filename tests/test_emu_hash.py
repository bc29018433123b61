"""CPU tests of csrc/hash.cuh (host build in tests/emu): SHA-256 against hashlib and HashG2WithDomain against the
reference's known answer (hash_test.go:72-82) and the host restatement pinned by it (bls_b200/hostmath.py)."""
import ctypes
import hashlib
import os
import sys

import numpy as np
import pytest

from bls_b200 import hostgen as hg, hostmath as hm, layout as L


@pytest.fixture(scope="module")
def emu():
    import __graft_entry__ as g
    return ctypes.CDLL(g.build_emu())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hash_cases(n, seed):
    rng = np.random.RandomState(seed)
    msgs = [bytes(32), b"\xff" * 32] + [rng.randint(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(n - 2)]
    doms = [bytes(8), b"\xff" * 8] + [rng.randint(0, 256, 8, dtype=np.uint8).tobytes() for _ in range(n - 2)]
    return msgs, doms


def expected_hashes(msgs, doms):
    return hg.g2_points([hm.hash_g2_with_domain(m, d) for m, d in zip(msgs, doms)])


def test_sha256_short(emu):
    rng = np.random.RandomState(3)
    for ln in (0, 1, 31, 41, 55):
        m = rng.randint(0, 256, max(ln, 1), dtype=np.uint8)
        dg = np.zeros(8, np.uint32)
        emu.emu_sha256_short(_p(m), ln, _p(dg))
        assert b"".join(int(w).to_bytes(4, "big") for w in dg) == hashlib.sha256(m[:ln].tobytes()).digest()


def test_hash_g2_with_domain_kat_and_host(emu, orc, kats):
    msgs, doms = hash_cases(6, 4)
    m = np.frombuffer(b"".join(msgs), np.uint8).copy(); d = np.frombuffer(b"".join(doms), np.uint8).copy()
    out = np.zeros(len(msgs), dtype=L.G2_AFFINE)
    emu.emu_hash_g2_with_domain(_p(m), _p(d), ctypes.c_size_t(1), ctypes.c_size_t(len(msgs)), _p(out))
    # hash_test.go:72-82: CompressG2(HashG2WithDomain(0^32, 0^8).ToAffine())
    assert orc.g2.compress(out[:1]).hex() == kats["hash"]["hash_g2_with_domain_zero_compressed"]
    assert out.tobytes() == expected_hashes(msgs, doms).tobytes()
    for i in range(len(msgs)):
        assert orc.g2.in_subgroup(out[i:i + 1])
    # one domain for the whole batch
    out1 = np.zeros(len(msgs), dtype=L.G2_AFFINE)
    emu.emu_hash_g2_with_domain(_p(m), _p(d), ctypes.c_size_t(0), ctypes.c_size_t(len(msgs)), _p(out1))
    assert out1.tobytes() == expected_hashes(msgs, [doms[0]] * len(msgs)).tobytes()
