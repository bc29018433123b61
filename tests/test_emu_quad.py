"""CPU unit tests of the four-lanes-per-pairing path (bls_b200/csrc/quad.cuh) through its host build: QL = 4, every lane
primitive loops over the four lanes of one quad, shuffles permute the lane array.  Every composite function the k_quad_*
kernels run is compared with the oracle (reference restatement) here, without a GPU."""
import ctypes

import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L

U64 = np.uint64


@pytest.fixture(scope="module", params=["sqr_dot2", "sqr_plain"])
def emu(request):
    """both forms of the lane Fq2 squaring (the kernels pick one by how full the machine is, quad.cuh::q2_sqr)"""
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_emu())
    lib.emu_set_lane_sqr_dot2(1 if request.param == "sqr_dot2" else 0)
    yield lib
    lib.emu_set_lane_sqr_dot2(1)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _op(emu, op, arg, a, b, o2=None):
    out = np.empty_like(a)
    ok = emu.emu_quad_fp12_op(op, ctypes.c_uint64(arg), _p(a), _p(b), _p(out), _p(o2), ctypes.c_size_t(a.shape[0]))
    return out, ok


def test_quad_fp12_ops(emu, orc):
    xs = orc.XorShift(112)
    n = 6
    a = xs.rand_fq(12 * n).reshape(n, 2, 3, 2, 6); b = xs.rand_fq(12 * n).reshape(n, 2, 3, 2, 6)
    a[1] = 0; a[1, 0, 0, 0] = L.fp_from_int(1)                       # the value 1
    a[2, 1] = 0                                                      # c1 = 0
    q1 = np.array(L.int_to_limbs(L.Q - 1), U64)
    a[3] = q1                                                        # every coefficient Q - 1
    for op, name, arg in [(0, "mul", 0), (3, "square", 0), (6, "inverse", 0), (7, "frobenius", 1), (7, "frobenius", 2),
                          (7, "frobenius", 3), (12, "conjugate", 0)]:
        out, _ = _op(emu, op, arg, a, b)
        assert (out == orc.fq12(name, a, b, arg)).all(), (name, arg)
    # in-place aliasing of the product: r = a * a through the general multiplication
    out, _ = _op(emu, 0, 0, a, a)
    assert (out == orc.fq12("square", a)).all()
    # sparse multiplication (fq12.go:32-47) with (c0, c1, c4) from b, and the extra product that rides in its free slot
    extra = np.zeros((n, 2, 6), U64)
    out, _ = _op(emu, 13, 0, a, b, extra)
    assert (out == orc.fq12("mul_by_014", a, b)).all()
    want = orc.fq2("mul", b[:, 1, 0], b[:, 1, 2])
    assert (extra == want).all()


def test_quad_inverse_of_zero(emu):
    z = np.zeros((1, 2, 3, 2, 6), U64)
    _, ok = _op(emu, 6, 0, z, z)
    assert ok == 0


def test_quad_cyclotomic(emu, orc):
    P = hg.g1_progression(3, 1, 3); Q = hg.g2_progression(4, 1, 3)
    f = orc.pairing_batch(P, Q)
    out, _ = _op(emu, 15, 0, f, f)
    assert (out == orc.fq12("square", f)).all()
    for x in (L.BLS_X, L.BLS_X >> 1):
        want = orc.fq12("conjugate", orc.fq12("exp", f, None, x))
        out, _ = _op(emu, 16, x, f, f)
        assert (out == want).all(), hex(x)
        out, _ = _op(emu, 17, x, f, f)
        assert (out == want).all(), hex(x)
    # the degenerate value 1 (compressed coordinates all zero): the warp-uniform fallback must give 1
    one = np.zeros_like(f[:1]); one[0, 0, 0, 0] = L.fp_from_int(1)
    out, _ = _op(emu, 16, L.BLS_X, one, one)
    assert (out == one).all()


def test_quad_miller_loop_and_final_exp(emu, orc, kats):
    P = np.concatenate([orc.g1_generator(), hg.g1_progression(0x99, 7, 4)])
    Q = np.concatenate([orc.g2_generator(), hg.g2_progression(0x55, 9, 4)])
    P["inf"][3] = 1
    n = P.size
    ml = np.zeros(n, dtype=L.FP12)
    emu.emu_quad_miller_loop(_p(P), _p(Q), ctypes.c_size_t(n), _p(ml))
    one = np.zeros((2, 3, 2, 6), U64); one[0, 0, 0] = L.fp_from_int(1)
    for i in range(n):
        want = one if P["inf"][i] else orc.miller_loop(P[i:i + 1], Q[i:i + 1])
        assert (ml[i] == want).all(), i
    fe = np.zeros(n, dtype=L.FP12); ok = np.zeros(n, np.uint8)
    emu.emu_quad_final_exp(_p(ml), ctypes.c_size_t(n), _p(fe), _p(ok))
    assert ok.all()
    P2 = P.copy(); P2["inf"][3] = 0
    want = orc.pairing_batch(P2, Q)
    for i in range(n):
        if P["inf"][i]:
            assert (fe[i].view(U64).reshape(2, 3, 2, 6) == one).all()
        else:
            assert fe[i].tobytes() == want[i].tobytes(), i
    exp = np.stack([L.fp_from_int(int(x, 16)) for x in kats["pairing_g1_g2"]["coeffs"]]).reshape(2, 3, 2, 6)
    assert (fe[0].view(U64).reshape(2, 3, 2, 6) == exp).all()        # RELIC vector, pairing_test.go:9-58


def test_quad_final_exp_zero_and_random(emu, orc):
    xs = orc.XorShift(33)
    f = xs.rand_fq(12 * 4).reshape(4, 2, 3, 2, 6)
    f[2] = 0
    fe = np.empty_like(f); ok = np.zeros(4, np.uint8)
    emu.emu_quad_final_exp(_p(f), ctypes.c_size_t(4), _p(fe), _p(ok))
    assert ok.tolist() == [1, 1, 0, 1]
    for i in (0, 1, 3):
        good, want = orc.final_exp(f[i])
        assert good and (fe[i] == want).all(), i


def test_quad_miller_loop_two_pairs(emu, orc):
    P = hg.g1_progression(0x31, 5, 6); Q = hg.g2_progression(0x47, 3, 6)
    P["inf"][4] = 1
    out = np.zeros(3, dtype=L.FP12)
    emu.emu_quad_miller_loop2(_p(P), _p(Q), ctypes.c_size_t(3), _p(out))
    for g in range(3):
        a = orc.miller_loop(P[2 * g:2 * g + 1], Q[2 * g:2 * g + 1]) if not P["inf"][2 * g] else None
        b = orc.miller_loop(P[2 * g + 1:2 * g + 2], Q[2 * g + 1:2 * g + 2])
        want = b if a is None else orc.fq12("mul", a, b)
        assert (out[g] == want).all(), g
