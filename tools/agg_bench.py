#!/usr/bin/env python3
"""Timing of the aggregation paths with device-resident inputs (CUDA events on the engine stream):
g1_sum, g1_msm, verify_aggregate_common_batch.  Points are a tiled 2^14-point progression (timing only)."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bls_b200 import capi, hostgen as hg, layout as L

ctx = capi.Ctx(0); st = torch.cuda.current_stream(); ctx.set_stream(st.cuda_stream)
dev = torch.device("cuda", 0)
def up(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
m = 1 << 14
base = hg.g1_progression(3, 5, m)
Kb, _ = hg.splitmix_scalars(1, m)
res = {}
sizes = [int(x) for x in sys.argv[1:]] or [1 << 16, 1 << 20, 1 << 22]
dO = torch.empty(288, dtype=torch.uint8, device=dev)
for n in sizes:
    dP = up(np.resize(base, n)); dK = up(np.resize(Kb, (n, 4)))
    res["g1_sum_%d_ms" % n] = timed(lambda: ctx.dev("b381_g1_sum_dev", dP.data_ptr(), ctypes.c_size_t(n), dO.data_ptr()))
    res["g1_msm_%d_ms" % n] = timed(lambda: ctx.dev("b381_g1_msm_dev", dP.data_ptr(), dK.data_ptr(), ctypes.c_size_t(n), dO.data_ptr()))
    print(json.dumps(res), flush=True)
# attestation batch: committees of 128 from a registry of 2^14
for natt in (1 << 12, 1 << 15):
    rng = np.random.RandomState(1)
    kidx = rng.randint(0, m, size=natt * 128).astype(np.uint32); koff = (np.arange(natt + 1) * 128).astype(np.uint32)
    H = hg.g2_progression(7, 11, 64); sig = np.resize(hg.g2_progression(9, 13, 256), natt)
    midx = rng.randint(0, 64, size=natt).astype(np.uint32)
    d = [up(x) for x in (base, kidx, koff, sig, H, midx)]
    dok = torch.empty(natt, dtype=torch.uint8, device=dev)
    res["verify_batch_%d_ms" % natt] = timed(lambda: ctx.dev("b381_verify_aggregate_common_batch_dev", *[x.data_ptr() for x in d], ctypes.c_size_t(natt), ctypes.c_size_t(base.size), ctypes.c_size_t(H.size), dok.data_ptr()), reps=2)
    print(json.dumps(res), flush=True)
