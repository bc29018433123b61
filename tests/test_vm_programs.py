"""The generated VM programs (bls_b200/vm/) executed by the big-integer emulator -- the same encoded
instruction stream the kernel interprets (csrc/vm.cuh) -- against the oracle: Miller loop, norm, final
exponentiation.  The emulator asserts the intermediate bounds the device arithmetic relies on."""
import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L
from bls_b200.vm import gen, sched as S, trace as T


@pytest.fixture(scope="module")
def progs():
    return {p["name"]: p for p in gen.build_all()}


def _ints(a):
    return [L.limbs_to_int(x) for x in np.asarray(a).reshape(-1, 6)]


def _run(pr, segs):
    return S.Emulator(pr["code"], pr["L"], pr["nslots"], pr["consts"]).run(segs)


def test_schedule_quality(progs):
    """the scheduler keeps the multiplier lanes busy and the working set inside shared memory"""
    for name in ("ml1", "fe_c"):
        st = progs[name]["stats"]
        assert st["mac_fill"] > 0.9, (name, st)
        assert progs[name]["nslots"] * 96 * (32 // progs[name]["L"]) <= 36 * 1024, name      # >= 6 warps per SM
    assert progs["ml1"]["stats"]["wide_macs"] <= 6916 * 300 * 1.01        # no more MACs than the thread-per-pairing Miller loop


@pytest.mark.parametrize("seed", [1, 2])
def test_pairing_pipeline_vs_oracle(progs, orc, seed):
    P = hg.g1_progression(0x99 + seed, 7, 1); Q = hg.g2_progression(0x55 + seed, 9, 1)
    if seed == 1:
        P, Q = orc.g1_generator(), orc.g2_generator()
    segs = _run(progs["ml1"], {0: _ints(P["x"]) + _ints(P["y"]), 1: _ints(Q["x"]) + _ints(Q["y"]), 2: [], 3: []})
    f = segs[2]
    assert f == _ints(orc.miller_loop(P, Q))                      # pre-final-exponentiation value is bit-identical too
    n = _run(progs["fe_a"], {0: f, 1: [], 2: [], 3: []})[2][0]
    ninv = pow(n * S.RINV % T.Q, -1, T.Q) * S.R % T.Q
    out = _run(progs["fe_c"], {0: f, 1: [ninv], 2: [], 3: []})[2]
    assert out == _ints(orc.pairing_batch(P, Q))


def test_final_exp_of_arbitrary_element(progs, orc):
    """a random Fq12 (not a Miller output): easy part, Frobenius tables and cyclotomic squarings on generic input"""
    f = orc.XorShift(9).rand_fq(12).reshape(1, 2, 3, 2, 6)
    fi = _ints(f)
    n = _run(progs["fe_a"], {0: fi, 1: [], 2: [], 3: []})[2][0]
    ninv = pow(n * S.RINV % T.Q, -1, T.Q) * S.R % T.Q
    out = _run(progs["fe_c"], {0: fi, 1: [ninv], 2: [], 3: []})[2]
    good, exp = orc.final_exp(f[0])
    assert good and out == _ints(exp)
