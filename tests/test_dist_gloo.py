"""The N>1 host logic on CPU: world-size-2 gloo process groups drive bls_b200/dist.py.  The
arithmetic stand-in is the host build of the device code (tests/emu) -- the same bucket-sharding
building blocks the kernels run -- and the expected value comes from the oracle."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class EmuEngine:
    """g1_msm_shard / g1_fold with the device headers compiled for the host (TEST-ONLY)"""

    def __init__(self):
        import __graft_entry__ as g
        from bls_b200 import layout as L
        self.L = L
        self.lib = ctypes.CDLL(g.build_emu())

    def g1_msm_shard(self, p, k, rank, nranks):
        out = np.zeros(1, dtype=self.L.G1_JAC)
        p = np.ascontiguousarray(p); k = np.ascontiguousarray(k)
        self.lib.emu_g1_msm(p.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.size),
                            0, rank, nranks, 16, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def g1_fold(self, parts):
        out = np.zeros(1, dtype=self.L.G1_JAC)
        parts = np.ascontiguousarray(parts)
        self.lib.emu_g1_fold(parts.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(parts.size), out.ctypes.data_as(ctypes.c_void_p))
        return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        from bls_b200 import dist as bd, hostgen as hg
        n = 96
        P = hg.g1_progression(0xD157, 3, n)
        K, _ = hg.splitmix_scalars(9, n)
        got = bd.msm_bucket_sharded(EmuEngine(), P, K)
        lo, hi = bd.tile(10, rank, world)
        flags = bd.gather_bytes(torch.full((3,), rank + 1, dtype=torch.uint8))
        ok = bd.all_ok(torch.tensor([1 if rank == 0 else 0], dtype=torch.int32))
        q.put((rank, got.tobytes(), (lo, hi), flags.tolist(), int(ok.item())))
    finally:
        dist.destroy_process_group()


def test_bucket_sharded_msm_world2(orc):
    from bls_b200 import hostgen as hg
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = hg.g1_progression(0xD157, 3, 96)
    K, _ = hg.splitmix_scalars(9, 96)
    exp = orc.g1.to_affine(orc.g1_msm_naive(P, K, threads=4)).tobytes()
    from bls_b200 import layout as L
    for rank, got, tile, flags, ok in res:
        assert orc.g1.to_affine(np.frombuffer(got, dtype=L.G1_JAC)).tobytes() == exp      # every rank holds the full sum
        assert flags == [1, 1, 1, 2, 2, 2] and ok == 0
    assert [r[2] for r in res] == [(0, 5), (5, 10)]


def test_tile_covers_everything():
    from bls_b200 import dist as bd
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [bd.tile(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
