#!/usr/bin/env python3
"""Latency of small batches through the C ABI (device-resident, CUDA events, best of 5): one CompareTwoPairings check and
pairing batches of 1 / 64 / 512 / 2048 units.  B381_VM_SPLIT=0 disables the two-lanes-per-operation VM form (A/B)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bls_b200 import capi, hostgen as hg

ctx = capi.Ctx(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


res = {"vm_split": os.environ.get("B381_VM_SPLIT", "1")}
a, b = 0x1234567, 0x89abcde
P = np.concatenate([hg.g1_mul(a), hg.g1_neg(hg.g1_mul(a * b))]); Q = np.concatenate([hg.g2_mul(b), hg.g2_mul(1)])
dP, dQ, dOff = up(P), up(Q), up(np.array([0, 2], np.uint32))
dOk = torch.zeros(8, dtype=torch.uint8, device=dev)
res["one_check_ms"] = timed(lambda: ctx.dev("b381_pairing_product_is_one_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(2), dOff.data_ptr(),
                                            ctypes.c_size_t(1), dOk.data_ptr()))
assert int(dOk[0].item()) == 1
for n in (1, 64, 512, 2048):
    Pn = np.resize(hg.g1_progression(3, 5, min(n, 64)), n); Qn = np.resize(hg.g2_progression(7, 11, min(n, 64)), n)
    dPn, dQn = up(Pn), up(Qn)
    dO = torch.empty(n * 576, dtype=torch.uint8, device=dev)
    res["pairing_batch_%d_ms" % n] = timed(lambda: ctx.dev("b381_pairing_batch_dev", dPn.data_ptr(), dQn.data_ptr(), ctypes.c_size_t(n), dO.data_ptr()))
print(json.dumps(res))
