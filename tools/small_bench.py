#!/usr/bin/env python3
"""Latency of small pairing batches (CUDA events, device-resident): VM vs thread-per-pairing kernels."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bls_b200 import capi, hostgen as hg
ctx = capi.Ctx(0); st = torch.cuda.current_stream(); ctx.set_stream(st.cuda_stream)
dev = torch.device("cuda", 0)
base_p = hg.g1_progression(3, 5, 512); base_q = hg.g2_progression(7, 11, 512)
res = {}
for n in (1, 2, 32, 512, 2048, 8192, 16384, 32768, 65536):
    P = np.resize(base_p, n); Q = np.resize(base_q, n)
    dP = torch.from_numpy(P.view(np.uint8).reshape(-1).copy()).to(dev); dQ = torch.from_numpy(Q.view(np.uint8).reshape(-1).copy()).to(dev)
    dO = torch.empty(n * 576, dtype=torch.uint8, device=dev)
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ctx.dev("b381_pairing_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(n), dO.data_ptr())
        e1.record(st); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    res[n] = round(best, 3)
print(os.environ.get("B381_VM", "1"), json.dumps(res))
