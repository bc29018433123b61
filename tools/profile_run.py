#!/usr/bin/env python3
"""Short driver for ncu captures: runs the pairing hot path (Miller-loop kernel + final-exp kernel)
on 2^16 pairs a few times.  Usage under gpurun, see profiles/README.md."""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bls_b200 import capi, hostgen as hg

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 16)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--what", default="pairing", choices=["pairing", "msm", "sum"])
a = ap.parse_args()

ctx = capi.Ctx(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda", 0)
m = min(a.n, 1024)
if a.what == "pairing":
    P = np.resize(hg.g1_progression(3, 5, m), a.n); Q = np.resize(hg.g2_progression(7, 11, m), a.n)
    dP = torch.from_numpy(P.view(np.uint8).reshape(-1).copy()).to(dev)
    dQ = torch.from_numpy(Q.view(np.uint8).reshape(-1).copy()).to(dev)
    dO = torch.empty(a.n * 576, dtype=torch.uint8, device=dev)
    for _ in range(a.reps):
        ctx.dev("b381_pairing_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(a.n), dO.data_ptr())
    torch.cuda.synchronize()
else:
    m = min(a.n, 1 << 14)
    P = np.resize(hg.g1_progression(3, 5, m), a.n)
    dP = torch.from_numpy(P.view(np.uint8).reshape(-1).copy()).to(dev)
    K, _ = hg.splitmix_scalars(1, min(a.n, 1 << 14))
    K = np.resize(K, (a.n, 4))
    dK = torch.from_numpy(K.view(np.uint8).reshape(-1).copy()).to(dev)
    dO = torch.empty(144, dtype=torch.uint8, device=dev)
    for _ in range(a.reps):
        if a.what == "msm":
            ctx.dev("b381_g1_msm_dev", dP.data_ptr(), dK.data_ptr(), ctypes.c_size_t(a.n), dO.data_ptr())
        else:
            ctx.dev("b381_g1_sum_dev", dP.data_ptr(), ctypes.c_size_t(a.n), dO.data_ptr())
    torch.cuda.synchronize()
print("done", a.what, a.n, "launches", ctx.launch_count)
