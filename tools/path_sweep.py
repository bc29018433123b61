#!/usr/bin/env python3
"""Pairing throughput of every kernel schedule (b381_set_kernel_path) over batch sizes: ms per b381_pairing_batch_dev call,
best of 3, resident inputs.  python tools/path_sweep.py [sizes...]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bls_b200 import capi, hostgen as hg
sizes = [int(x) for x in sys.argv[1:]] or [64, 592, 1024, 2048, 4096, 8192, 12288, 16384, 24576, 32768, 49152, 65536, 75776, 131072]
ctx = capi.Ctx(0); st = torch.cuda.current_stream(); ctx.set_stream(st.cuda_stream)
dev = torch.device("cuda", 0)
nmax = max(sizes); m = 2048
P = np.resize(hg.g1_progression(3, 5, m), nmax); Q = np.resize(hg.g2_progression(7, 11, m), nmax)
dP = torch.from_numpy(P.view(np.uint8).reshape(-1).copy()).to(dev); dQ = torch.from_numpy(Q.view(np.uint8).reshape(-1).copy()).to(dev)
dO = torch.empty(nmax * 576, dtype=torch.uint8, device=dev)
for n in sizes:
    row = {"n": n}
    for path in ("vm", "thread", "duo", "quad"):
        if path == "vm" and n > 32768:
            continue
        ctx.set_kernel_path(path)
        best = None
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            ctx.dev("b381_pairing_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(n), dO.data_ptr())
            e1.record(st); torch.cuda.synchronize()
            if rep:
                t = e0.elapsed_time(e1); best = t if best is None else min(best, t)
        row[path] = round(best, 3)
    row["best"] = min((v, k) for k, v in row.items() if k != "n")[1]
    print(json.dumps(row), flush=True)
