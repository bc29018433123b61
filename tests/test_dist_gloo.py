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


class EmuRlcEngine:
    """verify_rlc_partial / fp12_product_final_exp_is_one assembled from the host build of the device code (TEST-ONLY):
    the same arithmetic the kernels of csrc/b381.cu::verify_rlc_core run, one call at a time"""

    def __init__(self):
        import __graft_entry__ as g
        from bls_b200 import hostgen as hg, layout as L
        self.L, self.hg = L, hg
        self.lib = ctypes.CDLL(g.build_emu())

    def _p(self, a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def verify_rlc_partial(self, pub, h, sig, weights):
        L = self.L
        n = pub.size
        r = np.zeros((n, 4), np.uint64); r[:, 0] = np.asarray(weights, np.uint64)
        P = np.zeros(n + 1, dtype=L.G1_AFFINE); Q = np.zeros(n + 1, dtype=L.G2_AFFINE); rs = np.zeros(max(n, 1), dtype=L.G2_AFFINE)
        pub = np.ascontiguousarray(pub); sig = np.ascontiguousarray(sig)
        if n:
            self.lib.emu_g1_mul(self._p(pub), ctypes.c_size_t(1), self._p(r), ctypes.c_size_t(1), ctypes.c_size_t(n), self._p(P))
            self.lib.emu_g2_mul(self._p(sig), ctypes.c_size_t(1), self._p(r), ctypes.c_size_t(1), ctypes.c_size_t(n), self._p(rs))
            Q[:n] = h
        S = np.zeros(1, dtype=L.G2_JAC)
        self.lib.emu_g2_sum(self._p(rs), ctypes.c_size_t(n), self._p(S))
        P[n] = self.hg.g1_neg(self.hg.g1_mul(1))[0]
        Q[n]["x"] = S["x"][0]; Q[n]["y"] = S["y"][0]; Q[n]["inf"] = 0 if S["z"].any() else 1
        ml = np.zeros((n + 1, 72), np.uint64)
        self.lib.emu_miller_loop(self._p(P), self._p(Q), ctypes.c_size_t(n + 1), self._p(ml))
        return self._product(ml), int(not (pub["inf"].any() or sig["inf"].any()))

    def _product(self, vals):
        vals = np.ascontiguousarray(vals, np.uint64).reshape(-1, 72)
        acc = vals[:1].copy()
        for i in range(1, vals.shape[0]):
            out = np.zeros((1, 72), np.uint64)
            self.lib.emu_fp12_op(0, ctypes.c_uint64(0), self._p(acc), self._p(np.ascontiguousarray(vals[i:i + 1])), self._p(out), ctypes.c_size_t(1))
            acc = out
        return acc

    def fp12_product_final_exp_is_one(self, parts):
        prod = self._product(np.ascontiguousarray(parts))
        out = np.zeros((1, 72), np.uint64); ok = np.zeros(1, np.uint8)
        self.lib.emu_final_exp(self._p(prod), ctypes.c_size_t(1), self._p(out), self._p(ok))
        one = np.zeros((1, 72), np.uint64)
        one[0, :6] = self.L.fp_from_int(1)
        return bool(ok[0]) and out.tobytes() == one.tobytes()


def _rlc_worker(rank, world, port, q, corrupt):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bls_b200 import dist as bd, hostgen as hg
        n = 6                                             # triples in the whole job: sk_i = 11 + 3 i, H_i = (5 + i) G2
        sks = [11 + 3 * i for i in range(n)]; hs = [5 + i for i in range(n)]
        pubs = np.concatenate([hg.g1_mul(k) for k in sks]); H = np.concatenate([hg.g2_mul(h) for h in hs])
        sigs = np.concatenate([hg.g2_mul(k * h + (1 if (corrupt and i == 4) else 0)) for i, (k, h) in enumerate(zip(sks, hs))])
        w = np.array([(0x9E3779B97F4A7C15 ^ (i * 0x1234567)) | 1 for i in range(n)], dtype=np.uint64)
        lo, hi = bd.tile(n, rank, world)
        ok = bd.verify_rlc_sharded(EmuRlcEngine(), pubs[lo:hi], H[lo:hi], sigs[lo:hi], w[lo:hi])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _prebuild():
    """compile the host build of the device code ONCE in the parent, so that the spawned ranks only dlopen it"""
    import __graft_entry__ as g
    g.build_emu()


@pytest.mark.parametrize("corrupt", [False, True])
def test_rlc_verification_world2(corrupt):
    """SURVEY.md 8e, RLC variant: each rank's partial Miller product, one all-gather of 577 bytes, product + one final
    exponentiation on every rank; a single bad signature on rank 1 makes every rank answer False"""
    _prebuild()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 31500 + os.getpid() % 2000 + (7 if corrupt else 0)
    procs = [ctxm.Process(target=_rlc_worker, args=(r, 2, port, q, corrupt)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, not corrupt), (1, not corrupt)]


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
    _prebuild()
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
