#!/usr/bin/env python3
"""Does a 2^16 batch finish sooner when a slice of it runs on the two-lane kernels next to the one-pairing-per-thread kernels
(two contexts, two streams, concurrent kernels filling the thread slots the 2^16 batch leaves empty)?  Timing probe only.
usage: hybrid_probe.py [n_total] [slice sizes ...]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bls_b200 import capi, hostgen as hg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
slices = [int(x) for x in sys.argv[2:]] or [0, 4096, 8192, 10240, 12288, 16384]
dev = torch.device("cuda", 0)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
cA = capi.Ctx(0, path="thread"); cA.set_stream(sA.cuda_stream)
cB = capi.Ctx(0, path=os.environ.get("HYBRID_PATH", "duo")); cB.set_stream(sB.cuda_stream)
m = 2048
P = np.resize(hg.g1_progression(3, 5, m), n); Q = np.resize(hg.g2_progression(7, 11, m), n)
dP = torch.from_numpy(P.view(np.uint8).reshape(-1).copy()).to(dev); dQ = torch.from_numpy(Q.view(np.uint8).reshape(-1).copy()).to(dev)
dO = torch.empty(n * 576, dtype=torch.uint8, device=dev)
ref = None
for n2 in slices:
    n1 = n - n2
    best = None
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        sA.wait_event(e0); sB.wait_event(e0)
        if n1:
            cA.call("b381_pairing_batch_dev", ctypes.c_void_p(dP.data_ptr()), ctypes.c_void_p(dQ.data_ptr()), ctypes.c_size_t(n1), ctypes.c_void_p(dO.data_ptr()))
        if n2:
            cB.call("b381_pairing_batch_dev", ctypes.c_void_p(dP.data_ptr() + 104 * n1), ctypes.c_void_p(dQ.data_ptr() + 200 * n1), ctypes.c_size_t(n2),
                    ctypes.c_void_p(dO.data_ptr() + 576 * n1))
        ea, eb = torch.cuda.Event(), torch.cuda.Event()
        ea.record(sA); eb.record(sB)
        torch.cuda.current_stream().wait_event(ea); torch.cuda.current_stream().wait_event(eb)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        if rep: best = t if best is None else min(best, t)
    h = int(torch.sum(dO.view(torch.int64)[::977]).item())
    if ref is None: ref = h
    print(json.dumps({"n": n, "two_lane_slice": n2, "ms": round(best, 3), "pairings_per_s": round(n / best * 1e3), "same_output": h == ref}), flush=True)
