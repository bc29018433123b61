"""ctypes loader for libb381.so (include/b381.h) -- the only way Python reaches the engine.

There is no Python or CPU fallback: if the shared library is missing or no CUDA device is
usable, construction fails loudly.
"""
import ctypes
import os

import numpy as np

from . import layout as L

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B381_LIB") or os.path.join(_HERE, "libb381.so")   # B381_LIB: A/B builds of the same ABI

B381_OK = 0
ERRORS = {-1: "B381_ERR_ARG", -2: "B381_ERR_CUDA", -3: "B381_ERR_NOMEM", -4: "B381_ERR_NO_DEVICE"}


class B381Error(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("%s (%d) %s" % (ERRORS.get(code, "?"), code, msg))
        self.code = code


_lib = None


def load():
    """dlopen libb381.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s is not built; run __graft_entry__.build() (nvcc, sm_100a)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.b381_last_error.restype = ctypes.c_char_p
        lib.b381_launch_count.restype = ctypes.c_uint64
        _lib = lib
    return _lib


def _hp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Ctx:
    """one engine context bound to one CUDA device (b381_init / b381_free)"""

    PATHS = {"auto": -1, "thread": 0, "vm": 1, "quad": 2, "duo": 3}

    def __init__(self, device=0, path=None):
        self.lib = load()
        self._h = ctypes.c_void_p()
        rc = self.lib.b381_init(int(device), ctypes.byref(self._h))
        if rc != B381_OK:
            raise B381Error(rc, "b381_init(device=%d)" % device)
        self.device = device
        if path is not None:
            self.set_kernel_path(path)

    def set_kernel_path(self, path):
        """auto | thread | vm | quad | duo: which schedule of the pairing arithmetic is launched (b381_set_kernel_path)"""
        self.call("b381_set_kernel_path", ctypes.c_int(self.PATHS[path]))

    def close(self):
        if self._h:
            self.lib.b381_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != B381_OK:
            raise B381Error(rc, (self.lib.b381_last_error(self._h) or b"").decode())

    def call(self, name, *args):
        self._ck(getattr(self.lib, name)(self._h, *args))

    # -- test hook --------------------------------------------------------------------------
    _TEST_IN = {0: 6, 1: 12, 2: 36, 3: 72, 4: 72, 7: 72}

    def test_op(self, family, op, a, b=None, arg=0):
        """one field / tower / group-law operation of the device build (b381_test_op): returns (out, extra, ok)"""
        if family in (5, 6):
            dt, jdt = (L.G1_AFFINE, L.G1_JAC) if family == 5 else (L.G2_AFFINE, L.G2_JAC)
            a = np.ascontiguousarray(a, dtype=dt); b = a if b is None else np.ascontiguousarray(b, dtype=dt)
            n = a.size
            out = np.zeros(n, dtype=jdt)
        else:
            w = self._TEST_IN[family]
            a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, w)
            b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, w)
            n = a.shape[0]
            out = np.empty_like(a)
        assert (b.size if family in (5, 6) else b.shape[0]) == n
        extra = np.zeros((n, 12), np.uint64); ok = np.ones(n, np.uint8)
        self.call("b381_test_op", ctypes.c_int(family), ctypes.c_int(op), ctypes.c_uint64(arg), _hp(a), _hp(b), _hp(out), _hp(extra),
                  _hp(ok), ctypes.c_size_t(n))
        return out, extra, ok

    # -- plumbing ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        """launch on this cudaStream_t (0 = the legacy default stream, e.g. torch's default stream)"""
        self.call("b381_set_stream", ctypes.c_void_p(cuda_stream_handle or 0))

    def use_own_stream(self):
        self.call("b381_use_own_stream")

    def sync(self):
        self.call("b381_sync")

    @property
    def launch_count(self):
        return int(self.lib.b381_launch_count(self._h))

    # -- host-buffer entry points (numpy in / numpy out) ------------------------------------
    def pairing_batch(self, p, q):
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); q = np.ascontiguousarray(q, dtype=L.G2_AFFINE)
        assert p.size == q.size
        out = np.empty(p.size, dtype=L.FP12)
        self.call("b381_pairing_batch", _hp(p), _hp(q), ctypes.c_size_t(p.size), _hp(out))
        return out

    def pairing_batch_stream(self, p, q, batch):
        """the pairings of n pairs, `batch` per launch, copies overlapped with the kernels -- b381_pairing_batch_stream"""
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); q = np.ascontiguousarray(q, dtype=L.G2_AFFINE)
        assert p.size == q.size
        out = np.empty(p.size, dtype=L.FP12)
        self.call("b381_pairing_batch_stream", _hp(p), _hp(q), ctypes.c_size_t(p.size), ctypes.c_size_t(batch), _hp(out))
        return out

    def miller_loop_batch(self, p, q):
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); q = np.ascontiguousarray(q, dtype=L.G2_AFFINE)
        assert p.size == q.size
        out = np.empty(p.size, dtype=L.FP12)
        self.call("b381_miller_loop_batch", _hp(p), _hp(q), ctypes.c_size_t(p.size), _hp(out))
        return out

    def final_exp_batch(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64).reshape(-1, 2, 3, 2, 6)
        out = np.empty_like(f); ok = np.empty(f.shape[0], np.uint8)
        self.call("b381_final_exp_batch", _hp(f), ctypes.c_size_t(f.shape[0]), _hp(out), _hp(ok))
        return out, ok

    def g2_prepare_batch(self, q):
        """n x G2AffineToPrepared (g2.go:650-801) -- b381_g2_prepare_batch"""
        q = np.ascontiguousarray(q, dtype=L.G2_AFFINE)
        out = np.zeros(q.size, dtype=L.G2_PREPARED)
        self.call("b381_g2_prepare_batch", _hp(q), ctypes.c_size_t(q.size), _hp(out))
        return out

    def miller_loop_prepared_batch(self, p, prep, prep_idx=None):
        """MillerLoop of one item per pair with a prepared Q (pairing.go:4-7,16-75) -- b381_miller_loop_prepared_batch"""
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); prep = np.ascontiguousarray(prep, dtype=L.G2_PREPARED)
        idx = None if prep_idx is None else np.ascontiguousarray(prep_idx, dtype=np.uint32)
        out = np.empty(p.size, dtype=L.FP12)
        self.call("b381_miller_loop_prepared_batch", _hp(p), _hp(prep), ctypes.c_size_t(prep.size), _hp(idx), ctypes.c_size_t(p.size), _hp(out))
        return out

    def pairing_product_is_one(self, p, q, group_off):
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); q = np.ascontiguousarray(q, dtype=L.G2_AFFINE)
        off = np.ascontiguousarray(group_off, dtype=np.uint32)
        ok = np.empty(off.size - 1, np.uint8)
        self.call("b381_pairing_product_is_one", _hp(p), _hp(q), ctypes.c_size_t(p.size), _hp(off),
                  ctypes.c_size_t(off.size - 1), _hp(ok))
        return ok

    def g1_sum(self, p):
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); out = np.zeros(1, dtype=L.G1_JAC)
        self.call("b381_g1_sum", _hp(p), ctypes.c_size_t(p.size), _hp(out))
        return out

    def g2_sum(self, p):
        p = np.ascontiguousarray(p, dtype=L.G2_AFFINE); out = np.zeros(1, dtype=L.G2_JAC)
        self.call("b381_g2_sum", _hp(p), ctypes.c_size_t(p.size), _hp(out))
        return out

    def g1_msm(self, p, k):
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
        assert p.size == k.shape[0]
        out = np.zeros(1, dtype=L.G1_JAC)
        self.call("b381_g1_msm", _hp(p), _hp(k), ctypes.c_size_t(p.size), _hp(out))
        return out

    def g2_msm(self, p, k):
        p = np.ascontiguousarray(p, dtype=L.G2_AFFINE); k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
        assert p.size == k.shape[0]
        out = np.zeros(1, dtype=L.G2_JAC)
        self.call("b381_g2_msm", _hp(p), _hp(k), ctypes.c_size_t(p.size), _hp(out))
        return out

    def g1_msm_shard(self, p, k, rank, nranks):
        """this rank's bucket-sharded partial (Jacobian, not normalised) -- b381_g1_msm_shard_dev"""
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
        dp, dk, do = self.to_device(p), self.to_device(k), self.dev_empty(L.G1_JAC.itemsize)
        self.call("b381_g1_msm_shard_dev", dp.ptr, dk.ptr, ctypes.c_size_t(p.size), int(rank), int(nranks), do.ptr)
        return self.from_device(do, L.G1_JAC, 1)

    def g1_fold(self, parts):
        """normalised sum of Jacobian points -- b381_g1_fold_dev"""
        parts = np.ascontiguousarray(parts, dtype=L.G1_JAC)
        dp, do = self.to_device(parts), self.dev_empty(L.G1_JAC.itemsize)
        self.call("b381_g1_fold_dev", dp.ptr, ctypes.c_size_t(parts.size), do.ptr)
        return self.from_device(do, L.G1_JAC, 1)

    def verify_aggregate_common_batch(self, registry, key_idx, key_off, sig, msg_hash, msg_idx):
        """ok[a] for a batch of VerifyAggregateCommon checks -- b381_verify_aggregate_common_batch_dev"""
        registry = np.ascontiguousarray(registry, dtype=L.G1_AFFINE)
        key_idx = np.ascontiguousarray(key_idx, dtype=np.uint32); key_off = np.ascontiguousarray(key_off, dtype=np.uint32)
        sig = np.ascontiguousarray(sig, dtype=L.G2_AFFINE); msg_hash = np.ascontiguousarray(msg_hash, dtype=L.G2_AFFINE)
        msg_idx = np.ascontiguousarray(msg_idx, dtype=np.uint32)
        n = sig.size
        assert key_off.size == n + 1 and msg_idx.size == n
        bufs = [self.to_device(a) for a in (registry, key_idx, key_off, sig, msg_hash, msg_idx)]
        dok = self.dev_empty(max(n, 1))
        self.call("b381_verify_aggregate_common_batch_dev", *[b.ptr for b in bufs], ctypes.c_size_t(n), ctypes.c_size_t(registry.size),
                  ctypes.c_size_t(msg_hash.size), dok.ptr)
        return self.from_device(dok, np.uint8, n)

    def verify_aggregate_common_rlc(self, registry, key_idx, key_off, sig, msg_hash, msg_idx, weights=None):
        """ONE boolean for a batch of VerifyAggregateCommon checks (random linear combination grouped by message) --
        b381_verify_aggregate_common_rlc_dev.  weights: one non-zero 64-bit value per attestation; None draws them here."""
        registry = np.ascontiguousarray(registry, dtype=L.G1_AFFINE)
        key_idx = np.ascontiguousarray(key_idx, dtype=np.uint32); key_off = np.ascontiguousarray(key_off, dtype=np.uint32)
        sig = np.ascontiguousarray(sig, dtype=L.G2_AFFINE); msg_hash = np.ascontiguousarray(msg_hash, dtype=L.G2_AFFINE)
        msg_idx = np.ascontiguousarray(msg_idx, dtype=np.uint32)
        n = sig.size
        assert key_off.size == n + 1 and msg_idx.size == n
        weights = self.rlc_weights(n) if weights is None else np.asarray(weights, np.uint64)
        assert weights.size == n, "one 64-bit weight per attestation"
        self.set_rlc_weight_bits(64)
        r = np.zeros((n, 4), np.uint64); r[:, 0] = weights.reshape(-1)
        bufs = [self.to_device(a) for a in (registry, key_idx, key_off, sig, msg_hash, msg_idx, r)]
        dok = self.dev_empty(1)
        self.call("b381_verify_aggregate_common_rlc_dev", *[b.ptr for b in bufs], ctypes.c_size_t(n), ctypes.c_size_t(registry.size),
                  ctypes.c_size_t(msg_hash.size), dok.ptr)
        return bool(self.from_device(dok, np.uint8, 1)[0])

    # -- wire formats and scalar multiplication (csrc/codec.cuh) --------------------------------------------------
    def _decompress(self, name, data, nbytes, dtype, check_subgroup):
        raw = np.frombuffer(bytes(data), np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, np.uint8).reshape(-1)
        assert raw.size % nbytes == 0
        n = raw.size // nbytes
        out = np.zeros(n, dtype=dtype); status = np.zeros(n, np.uint8)
        self.call(name, _hp(raw), ctypes.c_size_t(n), int(bool(check_subgroup)), _hp(out), _hp(status))
        return out, status

    def g1_decompress_batch(self, data, check_subgroup=True):
        """n x 48 bytes -> (affine points, status codes): DecompressG1 / DecompressG1Unchecked"""
        return self._decompress("b381_g1_decompress_batch", data, 48, L.G1_AFFINE, check_subgroup)

    def g2_decompress_batch(self, data, check_subgroup=True):
        """n x 96 bytes -> (affine points, status codes): DecompressG2 / DecompressG2Unchecked"""
        return self._decompress("b381_g2_decompress_batch", data, 96, L.G2_AFFINE, check_subgroup)

    def g1_compress_batch(self, p):
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); out = np.zeros((p.size, 48), np.uint8)
        self.call("b381_g1_compress_batch", _hp(p), ctypes.c_size_t(p.size), _hp(out))
        return out

    def g2_compress_batch(self, p):
        p = np.ascontiguousarray(p, dtype=L.G2_AFFINE); out = np.zeros((p.size, 96), np.uint8)
        self.call("b381_g2_compress_batch", _hp(p), ctypes.c_size_t(p.size), _hp(out))
        return out

    def _mul(self, name, dtype, p, k):
        p = np.ascontiguousarray(p, dtype=dtype).reshape(-1); k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
        n = max(p.size, k.shape[0])
        assert p.size in (1, n) and k.shape[0] in (1, n)
        out = np.zeros(n, dtype=dtype)
        self.call(name, _hp(p), ctypes.c_size_t(0 if p.size == 1 and n > 1 else 1), _hp(k),
                  ctypes.c_size_t(0 if k.shape[0] == 1 and n > 1 else 1), ctypes.c_size_t(n), _hp(out))
        return out

    def g1_mul_batch(self, p, k):
        """out[i] = k[i] * p[i] (affine); a single point or a single scalar is broadcast -- MulFR + ToAffine"""
        return self._mul("b381_g1_mul_batch", L.G1_AFFINE, p, k)

    def g2_mul_batch(self, p, k):
        return self._mul("b381_g2_mul_batch", L.G2_AFFINE, p, k)

    def g1_mul_subgroup_batch(self, p, k):
        """the same product through the endomorphism ladder; every p[i] must lie in G1 (generator, hash, checked key)"""
        return self._mul("b381_g1_mul_subgroup_batch", L.G1_AFFINE, p, k)

    def g2_mul_subgroup_batch(self, p, k):
        return self._mul("b381_g2_mul_subgroup_batch", L.G2_AFFINE, p, k)

    def hash_g2_with_domain_batch(self, msgs32, domains8):
        """HashG2WithDomain over n 32-byte message hashes; domains8 is one 8-byte domain or n of them -> affine points"""
        m = np.frombuffer(b"".join(bytes(x) for x in msgs32), np.uint8) if not isinstance(msgs32, np.ndarray) else np.ascontiguousarray(msgs32, np.uint8).reshape(-1)
        d = np.frombuffer(bytes(domains8), np.uint8) if isinstance(domains8, (bytes, bytearray)) else \
            (np.ascontiguousarray(domains8, np.uint8).reshape(-1) if isinstance(domains8, np.ndarray) else np.frombuffer(b"".join(bytes(x) for x in domains8), np.uint8))
        assert m.size % 32 == 0 and d.size % 8 == 0
        n = m.size // 32
        assert d.size // 8 in (1, n)
        out = np.zeros(n, dtype=L.G2_AFFINE)
        self.call("b381_hash_g2_with_domain_batch", _hp(m.copy()), _hp(d.copy()), ctypes.c_size_t(0 if d.size == 8 and n > 1 else 1),
                  ctypes.c_size_t(n), _hp(out))
        return out

    def verify_with_domain_batch(self, pubs48, msgs32, domain8, sigs96):
        """ok[i] for n wire-format (public key, 32-byte message hash, signature) triples -- b381_verify_with_domain_batch"""
        cat = lambda xs, w: (np.ascontiguousarray(xs, np.uint8).reshape(-1) if isinstance(xs, np.ndarray)
                             else np.frombuffer(b"".join(bytes(x) for x in xs), np.uint8).copy())
        p, m, s = cat(pubs48, 48), cat(msgs32, 32), cat(sigs96, 96)
        d = np.frombuffer(bytes(domain8), np.uint8).copy() if isinstance(domain8, (bytes, bytearray)) else cat(domain8, 8)
        n = p.size // 48
        assert p.size == 48 * n and m.size == 32 * n and s.size == 96 * n and d.size in (8, 8 * n)
        ok = np.zeros(n, np.uint8)
        self.call("b381_verify_with_domain_batch", _hp(p), _hp(m), _hp(d), ctypes.c_size_t(0 if d.size == 8 and n > 1 else 1), _hp(s),
                  ctypes.c_size_t(n), _hp(ok))
        return ok

    @staticmethod
    def _pack(msgs):
        off = np.zeros(len(msgs) + 1, np.uint64)
        if len(msgs):
            off[1:] = np.cumsum([len(m) for m in msgs])
        raw = np.frombuffer(b"".join(bytes(m) for m in msgs) + b"\0", np.uint8).copy()
        return raw, off

    def hash_g1_batch(self, msgs):
        """HashG1 of every message (bytes) -> affine G1 points -- b381_hash_g1_batch"""
        raw, off = self._pack(msgs); out = np.zeros(len(msgs), dtype=L.G1_AFFINE)
        self.call("b381_hash_g1_batch", _hp(raw), _hp(off), ctypes.c_size_t(len(msgs)), _hp(out))
        return out

    def hash_g2_batch(self, msgs):
        """HashG2 of every message (bytes) -> affine G2 points -- b381_hash_g2_batch"""
        raw, off = self._pack(msgs); out = np.zeros(len(msgs), dtype=L.G2_AFFINE)
        self.call("b381_hash_g2_batch", _hp(raw), _hp(off), ctypes.c_size_t(len(msgs)), _hp(out))
        return out

    def _verify_wire(self, name, pubs, pb, msgs, sigs, sb):
        p = np.ascontiguousarray(pubs, np.uint8).reshape(-1); s = np.ascontiguousarray(sigs, np.uint8).reshape(-1)
        n = len(msgs)
        assert p.size == pb * n and s.size == sb * n
        raw, off = self._pack(msgs); ok = np.zeros(n, np.uint8)
        self.call(name, _hp(p), _hp(raw), _hp(off), _hp(s), ctypes.c_size_t(n), _hp(ok))
        return ok

    def g1pubs_verify_batch(self, pubs48, msgs, sigs96):
        """ok[i] = g1pubs.Verify(msgs[i], pub_i, sig_i) from wire bytes -- b381_g1pubs_verify_batch"""
        return self._verify_wire("b381_g1pubs_verify_batch", pubs48, 48, msgs, sigs96, 96)

    def g2pubs_verify_batch(self, pubs96, msgs, sigs48):
        """ok[i] = g2pubs.Verify(msgs[i], pub_i, sig_i) from wire bytes -- b381_g2pubs_verify_batch"""
        return self._verify_wire("b381_g2pubs_verify_batch", pubs96, 96, msgs, sigs48, 48)

    @staticmethod
    def rlc_weights(n):
        """n fresh non-zero 64-bit weights from the operating system's generator (the verifier's randomness)"""
        w = np.frombuffer(os.urandom(8 * n), dtype=np.uint64).copy()
        w[w == 0] = 1
        return w

    def set_rlc_weight_bits(self, bits):
        self.call("b381_set_rlc_weight_bits", ctypes.c_int(bits))

    def verify_with_domain_rlc_batch(self, pubs48, msgs32, domain8, sigs96, weights=None):
        """one boolean for the whole batch (random linear combination) -- b381_verify_with_domain_rlc_batch.  weights: one
        non-zero 64-bit value per triple; None draws them here (a zero weight is rejected by the engine: the check is false)"""
        p = np.ascontiguousarray(pubs48, np.uint8).reshape(-1); m = np.ascontiguousarray(msgs32, np.uint8).reshape(-1)
        s = np.ascontiguousarray(sigs96, np.uint8).reshape(-1); d = np.frombuffer(bytes(domain8), np.uint8).copy()
        n = p.size // 48
        assert p.size == 48 * n and s.size == 96 * n and m.size == 32 * n and d.size == 8, "buffer sizes do not match the triple count"
        weights = self.rlc_weights(n) if weights is None else np.asarray(weights, np.uint64)
        assert weights.size == n
        self.set_rlc_weight_bits(64)
        r = np.zeros((n, 4), np.uint64); r[:, 0] = weights
        ok = np.zeros(1, np.uint8)
        self.call("b381_verify_with_domain_rlc_batch", _hp(p), _hp(m), _hp(d), ctypes.c_size_t(0), _hp(s), _hp(r), ctypes.c_size_t(n), _hp(ok))
        return bool(ok[0])

    # -- multi-GPU random-linear-combination check: per-rank partial products and the finishing step ----------------
    def verify_rlc_partial(self, pub, msg_point, sig, weights):
        """(partial Fq12 product as a 1-element FP12 array, valid flag) for this rank's triples -- b381_verify_rlc_partial_dev"""
        pub = np.ascontiguousarray(pub, dtype=L.G1_AFFINE); h = np.ascontiguousarray(msg_point, dtype=L.G2_AFFINE)
        sig = np.ascontiguousarray(sig, dtype=L.G2_AFFINE)
        n = pub.size
        assert h.size == n and sig.size == n and np.asarray(weights).size == n
        self.set_rlc_weight_bits(64)
        r = np.zeros((n, 4), np.uint64); r[:, 0] = np.asarray(weights, np.uint64)
        bufs = [self.to_device(a) for a in (pub, h, sig, r)]
        dpart, dval = self.dev_empty(576), self.dev_empty(8)
        self.call("b381_verify_rlc_partial_dev", *[b.ptr for b in bufs], ctypes.c_size_t(n), dpart.ptr, dval.ptr)
        return self.from_device(dpart, np.uint64, 72).reshape(1, 72), int(self.from_device(dval, np.uint8, 1)[0])

    def fp12_product_final_exp_is_one(self, parts):
        """FinalExponentiation(prod parts) == 1 -- b381_fp12_product_final_exp_is_one_dev"""
        parts = np.ascontiguousarray(parts, dtype=np.uint64).reshape(-1, 72)
        dp, dok = self.to_device(parts), self.dev_empty(8)
        self.call("b381_fp12_product_final_exp_is_one_dev", dp.ptr, ctypes.c_size_t(parts.shape[0]), dok.ptr)
        return bool(self.from_device(dok, np.uint8, 1)[0])

    def miller_product(self, p, q):
        """prod_i MillerLoop(p[i], q[i]) without the final exponentiation -- b381_miller_product_dev"""
        p = np.ascontiguousarray(p, dtype=L.G1_AFFINE); q = np.ascontiguousarray(q, dtype=L.G2_AFFINE)
        dp, dq, do = self.to_device(p), self.to_device(q), self.dev_empty(576)
        self.call("b381_miller_product_dev", dp.ptr, dq.ptr, ctypes.c_size_t(p.size), do.ptr)
        return self.from_device(do, np.uint64, 72).reshape(1, 72)

    def verify_aggregate_common_with_domain_batch(self, registry, key_idx, key_off, sigs96, msgs32, domain8, msg_idx):
        """ok[a] for attestations given with compressed signatures and 32-byte message hashes --
        b381_verify_aggregate_common_with_domain_batch_dev"""
        registry = np.ascontiguousarray(registry, dtype=L.G1_AFFINE)
        key_idx = np.ascontiguousarray(key_idx, dtype=np.uint32); key_off = np.ascontiguousarray(key_off, dtype=np.uint32)
        sig = np.ascontiguousarray(sigs96, np.uint8).reshape(-1, 96); msg = np.ascontiguousarray(msgs32, np.uint8).reshape(-1, 32)
        dom = np.frombuffer(bytes(domain8), np.uint8).copy(); msg_idx = np.ascontiguousarray(msg_idx, dtype=np.uint32)
        n = sig.shape[0]
        assert key_off.size == n + 1 and msg_idx.size == n and dom.size == 8
        b = [self.to_device(a) for a in (registry, key_idx, key_off, sig, msg, dom, msg_idx)]
        dok = self.dev_empty(max(n, 1))
        self.call("b381_verify_aggregate_common_with_domain_batch_dev", b[0].ptr, b[1].ptr, b[2].ptr, b[3].ptr, b[4].ptr,
                  ctypes.c_size_t(msg.shape[0]), b[5].ptr, b[6].ptr, ctypes.c_size_t(n), ctypes.c_size_t(registry.size), dok.ptr)
        return self.from_device(dok, np.uint8, n)

    # -- raw device buffers owned by the engine (b381_dev_alloc / b381_h2d / b381_d2h) ----------
    def dev_empty(self, nbytes):
        return DevBuf(self, nbytes)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        b = DevBuf(self, arr.nbytes)
        if arr.nbytes:
            self.call("b381_h2d", b.ptr, _hp(arr), ctypes.c_size_t(arr.nbytes))
        b.keep = arr          # the copy is asynchronous: keep the source alive with the buffer
        return b

    def from_device(self, buf, dtype, count):
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            self.call("b381_d2h", _hp(out), buf.ptr, ctypes.c_size_t(out.nbytes))   # synchronises the stream
        else:
            self.sync()
        return out

    # -- device-pointer entry points (ints are raw device addresses, e.g. torch .data_ptr()) --
    def dev(self, name, *args):
        conv = [ctypes.c_void_p(a) if isinstance(a, int) and not isinstance(a, bool) and a > 0xFFFF else a for a in args]
        self.call(name, *conv)


class DevBuf:
    """device allocation made through the C ABI (b381_dev_alloc); freed with the object"""

    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        self.ptr = ctypes.c_void_p()
        self.keep = None
        ctx.call("b381_dev_alloc", ctypes.c_size_t(max(self.nbytes, 1)), ctypes.byref(self.ptr))

    def __del__(self):
        try:
            if self.ptr and self.ctx._h:
                self.ctx.lib.b381_dev_free(self.ctx._h, self.ptr)
        except Exception:
            pass
