"""The VM interpreter kernel (bls_b200/csrc/vm.cuh) against the big-integer emulator on RANDOM programs:
every operation shape (products, squares, sums; xi / conj / negative terms; tripled groups; Fq and Fq2
loads and stores) on random field elements including the edge values 0, 1, Q-1."""
import ctypes
import random

import numpy as np
import pytest

from bls_b200 import layout as L
from bls_b200.vm import sched as S, trace as T

pytestmark = pytest.mark.gpu
LANES = 4


@pytest.fixture(scope="module")
def ctx():
    from bls_b200 import capi
    c = capi.Ctx(0)
    yield c
    c.close()


def random_program(rng, nin, nops):
    """ops straight in the folded form: random operands / addends over previously defined values"""
    p = T.Program("rnd")
    vals = [p.load(0, 2 * i).single()[0] for i in range(nin)]
    ops = list(p.ops)

    def term():
        return (rng.choice(vals), rng.choice((1, -1)), rng.choice((0, 0, 1)), rng.choice((0, 0, 0, 1)))
    outs = []
    for _ in range(nops):
        mode = rng.choice(("mul", "mul", "sqr", "lin"))
        d = p._new()
        # operand shapes the tracer emits: squares of single values, products with na * nb <= 4
        na, nb = rng.choice(((1, 1), (1, 2), (2, 1), (2, 2), (1, 3), (3, 1))) if mode == "mul" else (1, 0)
        a = [term() for _ in range(na)] if mode != "lin" else []
        b = [term() for _ in range(nb)] if mode == "mul" else []
        if mode == "sqr" and rng.random() < 0.7:
            a = [(a[0][0], 1, 0, 0)]                    # the plain single-term fast path
        m3 = rng.random() < 0.3
        g1 = [term() for _ in range(rng.randint(0, 4))] if m3 else []
        add = [term() for _ in range(rng.randint(0 if (a or g1) else 1, 8 - len(g1)))]
        ops.append({"kind": "FMA", "mode": mode, "dst": d, "a": a, "b": b, "g1": g1, "m3": 1 if m3 else 0, "add": add,
                    "pxi": rng.choice((0, 1)) if mode != "lin" else 0})
        vals.append(d); outs.append(d)
    for i, v in enumerate(outs):
        ops.append({"kind": "IO", "op": "store", "src": v, "seg": 2, "idx": 2 * i, "width": 2, "dst_handle": -1000 - i})
    return ops, len(outs)


@pytest.mark.parametrize("seed", range(6))
def test_random_programs_match_emulator(ctx, seed):
    rng = random.Random(seed)
    nin, nops, nunits = 5, 40, 19
    ops, nout = random_program(rng, nin, nops)
    steps = S.schedule(ops, LANES, 0)
    slot, nslots = S.allocate(ops, steps)
    code = S.encode(ops, steps, slot, LANES)
    edge = [0, 1, T.Q - 1, T.Q - 2, 2]
    ins = np.zeros((nunits, 2 * nin, 6), np.uint64)
    ints = []
    for u in range(nunits):
        row = [rng.choice(edge) if rng.random() < 0.25 else rng.randrange(T.Q) for _ in range(2 * nin)]
        ints.append(row)
        for j, v in enumerate(row):
            ins[u, j] = np.array(L.int_to_limbs(v), np.uint64)
    d_in = ctx.to_device(ins)
    d_out = ctx.dev_empty(nunits * 2 * nout * 48)
    d_x = ctx.dev_empty(64)
    segs = (ctypes.c_void_p * 4)(d_in.ptr.value, d_x.ptr.value, d_out.ptr.value, d_x.ptr.value)
    strides = (ctypes.c_size_t * 4)(2 * nin * 48, 0, 2 * nout * 48, 0)
    ctx.call("b381_vm_exec_dev", ctypes.c_char_p(code), LANES, len(steps), nslots, None, 0, segs, strides, ctypes.c_size_t(nunits))
    got = ctx.from_device(d_out, np.uint64, nunits * 2 * nout * 6).reshape(nunits, 2 * nout, 6)
    emu = S.Emulator(code, LANES, nslots, [])
    for u in range(nunits):
        exp = emu.run({0: list(ints[u]), 1: [], 2: [], 3: []})[2]
        gotu = [L.limbs_to_int(x) for x in got[u]]
        assert gotu == exp, (seed, u, [i for i in range(len(exp)) if gotu[i] != exp[i]][:4])
