"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/oracle.hpp).  The product (bls_b200/) never does.
"""
import ctypes
import os
import subprocess
import numpy as np

from bls_b200 import layout as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".hpp", ".inc"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_fq_mul_count.restype = ctypes.c_uint64
        _lib.orc_time_pairings.restype = ctypes.c_double
        for n in ("orc_mac_with_carry", "orc_add_with_carry", "orc_sub_with_borrow"):
            getattr(_lib, n).restype = ctypes.c_uint64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


U64 = np.uint64
FP = np.dtype((U64, (6,)))
FP2 = np.dtype((U64, (2, 6)))
FP6 = np.dtype((U64, (3, 2, 6)))

FQ_OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, double=5, inverse=6, from_repr=7, to_repr=8, sqrt=9)
FQ2_OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, double=5, inverse=6, frobenius=7, mul_by_nonresidue=8, sqrt=9)
FQ6_OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, inverse=6, frobenius=7, mul_by_nonresidue=8, mul_by_1=10, mul_by_01=11)
FQ12_OPS = dict(mul=0, square=3, inverse=6, frobenius=7, conjugate=12, mul_by_014=13, exp=14)


def _binop(fn, shape_tail, op, a, b, *extra):
    a = np.ascontiguousarray(a, dtype=U64).reshape((-1,) + shape_tail)
    b = a if b is None else np.ascontiguousarray(b, dtype=U64).reshape((-1,) + shape_tail)
    out = np.empty_like(a)
    fn(op, *extra, _p(a), _p(b), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def fq(op, a, b=None):
    return _binop(lib().orc_fq_op, (6,), FQ_OPS[op], a, b)


def fq2(op, a, b=None):
    return _binop(lib().orc_fq2_op, (2, 6), FQ2_OPS[op], a, b)


def fq6(op, a, b=None, arg=0):
    return _binop(lib().orc_fq6_op, (3, 2, 6), FQ6_OPS[op], a, b, ctypes.c_int(arg))


def fq12(op, a, b=None, arg=0):
    return _binop(lib().orc_fq12_op, (2, 3, 2, 6), FQ12_OPS[op], a, b, ctypes.c_uint64(arg))


def g1_generator():
    o = np.zeros(1, dtype=L.G1_AFFINE)
    lib().orc_g1_generator(_p(o))
    return o


def g2_generator():
    o = np.zeros(1, dtype=L.G2_AFFINE)
    lib().orc_g2_generator(_p(o))
    return o


def frobenius_tables():
    a = np.zeros((6, 2, 6), U64); b = np.zeros((6, 2, 6), U64); c = np.zeros((12, 2, 6), U64)
    lib().orc_frobenius_tables(_p(a), _p(b), _p(c))
    return a, b, c


def _grp(prefix, AFF, JAC):
    class G:
        @staticmethod
        def add(a, b):
            a = _c(a, JAC); b = _c(b, JAC); o = np.empty_like(a)
            getattr(lib(), prefix + "_add")(_p(a), _p(b), _p(o), ctypes.c_size_t(a.size)); return o

        @staticmethod
        def add_affine(a, b):
            a = _c(a, JAC); b = _c(b, AFF); o = np.empty_like(a)
            getattr(lib(), prefix + "_add_affine")(_p(a), _p(b), _p(o), ctypes.c_size_t(a.size)); return o

        @staticmethod
        def double(a):
            a = _c(a, JAC); o = np.empty_like(a)
            getattr(lib(), prefix + "_double")(_p(a), _p(o), ctypes.c_size_t(a.size)); return o

        @staticmethod
        def to_affine(a):
            a = _c(a, JAC); o = np.zeros(a.shape, dtype=AFF)
            getattr(lib(), prefix + "_to_affine")(_p(a), _p(o), ctypes.c_size_t(a.size)); return o

        @staticmethod
        def to_proj(a):
            a = _c(a, AFF); o = np.zeros(a.shape, dtype=JAC)
            getattr(lib(), prefix + "_to_proj")(_p(a), _p(o), ctypes.c_size_t(a.size)); return o

        @staticmethod
        def mul_fr(a, s, threads=1):
            a = _c(a, AFF); s = np.ascontiguousarray(s, dtype=U64).reshape(-1, 4); o = np.zeros(a.shape, dtype=JAC)
            getattr(lib(), prefix + "_affine_mul_fr")(_p(a), _p(s), _p(o), ctypes.c_size_t(a.size), threads); return o

        @staticmethod
        def sum_proj(a):
            a = _c(a, JAC); o = np.zeros(1, dtype=JAC)
            getattr(lib(), prefix + "_sum_proj")(_p(a), ctypes.c_size_t(a.size), _p(o)); return o

        @staticmethod
        def sum_affine(a):
            a = _c(a, AFF); o = np.zeros(1, dtype=JAC)
            getattr(lib(), prefix + "_sum_affine")(_p(a), ctypes.c_size_t(a.size), _p(o)); return o

        @staticmethod
        def proj_equal(a, b):
            a = _c(a, JAC); b = _c(b, JAC)
            return bool(getattr(lib(), prefix + "_proj_equal")(_p(a), _p(b)))

        @staticmethod
        def is_on_curve(a):
            a = _c(a, AFF); return bool(getattr(lib(), prefix + "_is_on_curve")(_p(a)))

        @staticmethod
        def in_subgroup(a):
            a = _c(a, AFF); return bool(getattr(lib(), prefix + "_in_subgroup")(_p(a)))

        @staticmethod
        def compress(a):
            a = _c(a, AFF); o = np.zeros(48 if AFF is L.G1_AFFINE else 96, np.uint8)
            getattr(lib(), prefix + "_compress")(_p(a), _p(o)); return o.tobytes()

        @staticmethod
        def decompress(b, checked=True):
            i = np.frombuffer(bytes(b), np.uint8).copy(); o = np.zeros(1, dtype=AFF)
            e = getattr(lib(), prefix + "_decompress")(_p(i), _p(o), int(checked)); return e, o
    return G


g1 = _grp("orc_g1", L.G1_AFFINE, L.G1_JAC)
g2 = _grp("orc_g2", L.G2_AFFINE, L.G2_JAC)


def g1_msm_naive(points, scalars, threads=1):
    a = _c(points, L.G1_AFFINE); s = np.ascontiguousarray(scalars, dtype=U64).reshape(-1, 4); o = np.zeros(1, dtype=L.G1_JAC)
    lib().orc_g1_msm_naive(_p(a), _p(s), ctypes.c_size_t(a.size), _p(o), threads)
    return o


def g2_prepare(q):
    q = _c(q, L.G2_AFFINE); co = np.zeros((68, 3, 2, 6), U64)
    n = lib().orc_g2_prepare(_p(q), _p(co))
    return co[:n]


def miller_loop(p, q):
    p = _c(p, L.G1_AFFINE); q = _c(q, L.G2_AFFINE); o = np.zeros(1, dtype=L.FP12)
    lib().orc_miller_loop(_p(p), _p(q), ctypes.c_size_t(p.size), _p(o))
    return o[0]


def final_exp(f):
    f = np.ascontiguousarray(f, dtype=U64).reshape(1, 2, 3, 2, 6); o = np.zeros_like(f)
    ok = lib().orc_final_exp(_p(f), _p(o))
    return bool(ok), o[0]


def pairing_batch(p, q, threads=1):
    p = _c(p, L.G1_AFFINE); q = _c(q, L.G2_AFFINE); o = np.zeros(p.size, dtype=L.FP12)
    lib().orc_pairing_batch(_p(p), _p(q), ctypes.c_size_t(p.size), _p(o), threads)
    return o


def pairing_product_is_one(p, q, offsets, threads=1):
    p = _c(p, L.G1_AFFINE); q = _c(q, L.G2_AFFINE); off = np.ascontiguousarray(offsets, dtype=np.uint32)
    ok = np.zeros(off.size - 1, np.uint8)
    lib().orc_pairing_product_is_one(_p(p), _p(q), _p(off), ctypes.c_size_t(off.size - 1), _p(ok), threads)
    return ok


def compare_two_pairings(p1, q1, p2, q2):
    return bool(lib().orc_compare_two_pairings(_p(_c(p1, L.G1_JAC)), _p(_c(q1, L.G2_JAC)), _p(_c(p2, L.G1_JAC)), _p(_c(q2, L.G2_JAC))))


class XorShift:
    """the tests' deterministic io.Reader (g1_test.go:106-124) feeding Go's crypto/rand.Int"""

    def __init__(self, seed):
        self.state = ctypes.c_uint64(seed)

    def rand_fr(self, n=1):
        o = np.zeros((n, 4), U64)
        lib().orc_rand_fr(ctypes.byref(self.state), _p(o), ctypes.c_size_t(n)); return o

    def rand_fq(self, n=1):
        o = np.zeros((n, 6), U64)
        lib().orc_rand_fq(ctypes.byref(self.state), _p(o), ctypes.c_size_t(n)); return o


def fq_mul_count(reset=False):
    if reset:
        lib().orc_fq_mul_count_reset()
    return int(lib().orc_fq_mul_count())


def time_pairings(p, q, n, threads):
    p = _c(p, L.G1_AFFINE); q = _c(q, L.G2_AFFINE); last = np.zeros(1, dtype=L.FP12)
    return float(lib().orc_time_pairings(_p(p), _p(q), ctypes.c_size_t(p.size), ctypes.c_size_t(n), threads, _p(last)))
