"""GPU parity of b381_hash_g2_with_domain_batch (HashG2WithDomain, g2.go:1041-1085) through the C ABI: the reference's
known answer (hash_test.go:72-82), the pinned host restatement on random inputs, and -- at batch scale -- the
size-independent properties (every output is on the curve and in the r-torsion; equal inputs give equal outputs)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_emu_hash import hash_cases, expected_hashes

from bls_b200 import layout as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from bls_b200 import capi
    return capi.Ctx(0)


def test_hash_g2_with_domain_matches_reference(ctx, orc, kats):
    msgs, doms = hash_cases(24, 5)
    out = ctx.hash_g2_with_domain_batch(msgs, doms)
    assert orc.g2.compress(out[:1]).hex() == kats["hash"]["hash_g2_with_domain_zero_compressed"]
    assert out.tobytes() == expected_hashes(msgs, doms).tobytes()
    out1 = ctx.hash_g2_with_domain_batch(msgs, doms[3])
    assert out1.tobytes() == expected_hashes(msgs, [doms[3]] * len(msgs)).tobytes()


def test_hash_batch_properties(ctx, orc):
    n = 4096
    rng = np.random.RandomState(6)
    m = rng.randint(0, 256, (n, 32), dtype=np.uint8)
    m[n // 2:] = m[:n // 2]                                   # duplicates
    out = ctx.hash_g2_with_domain_batch(m, bytes([1, 2, 3, 4, 5, 6, 7, 8]))
    assert out[:n // 2].tobytes() == out[n // 2:].tobytes()
    assert not out["inf"].any()
    for i in range(0, n // 2, 97):
        assert orc.g2.is_on_curve(out[i:i + 1]) and orc.g2.in_subgroup(out[i:i + 1])
    # the decompression kernel agrees that they are valid signatures-to-be: compress -> checked decompress round trip
    back, st = ctx.g2_decompress_batch(ctx.g2_compress_batch(out).tobytes(), check_subgroup=True)
    assert not st.any() and back.tobytes() == out.tobytes()
