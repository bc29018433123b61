#!/usr/bin/env python3
"""Quick A/B timing of the pairing kernels (no CPU baseline, tiled inputs): prints ms per kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bls_b200 import capi, hostgen as hg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
ctx = capi.Ctx(0); st = torch.cuda.current_stream(); ctx.set_stream(st.cuda_stream)
dev = torch.device("cuda", 0)
m = min(n, 2048)
P = np.resize(hg.g1_progression(3, 5, m), n); Q = np.resize(hg.g2_progression(7, 11, m), n)
dP = torch.from_numpy(P.view(np.uint8).reshape(-1).copy()).to(dev); dQ = torch.from_numpy(Q.view(np.uint8).reshape(-1).copy()).to(dev)
dM = torch.empty(n * 576, dtype=torch.uint8, device=dev); dO = torch.empty(n * 576, dtype=torch.uint8, device=dev)
res = {}
for rep in range(4):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(st)
    ctx.dev("b381_miller_loop_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(n), dM.data_ptr())
    e[1].record(st)
    ctx.dev("b381_final_exp_batch_dev", dM.data_ptr(), ctypes.c_size_t(n), dO.data_ptr(), None)
    e[2].record(st)
    torch.cuda.synchronize()
    res = {"miller_ms": e[0].elapsed_time(e[1]), "final_exp_ms": e[1].elapsed_time(e[2])}
res["pairings_per_s"] = n / ((res["miller_ms"] + res["final_exp_ms"]) * 1e-3)
print(os.environ.get("B381_LIB", "default"), n, res)
