#!/usr/bin/env python3
"""One bucket-sharded G1 MSM call per (rank, nranks) given on the command line, on tiled points and splitmix scalars
(timing only): python tools/msm_phases.py LOG2N rank:nranks ...   Run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel times of each call, or plain for CUDA-event totals."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bls_b200 import capi, hostgen as hg, layout as L
lg = int(sys.argv[1]); n = 1 << lg
ctx = capi.Ctx(0); st = torch.cuda.current_stream(); ctx.set_stream(st.cuda_stream)
dev = torch.device("cuda", 0)
m = 1 << 12
P = np.resize(hg.g1_progression(3, 5, m), n)
DISTINCT = os.environ.get("MSM_DISTINCT") == "1"     # n different points (i + 1) G made on the device: the gathers of the chunk sums miss the caches
rng = np.random.RandomState(7)
K = rng.randint(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
K[:, 3] &= np.uint64((1 << 62) - 1)                      # < 2^254 < r: canonical
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)
dP, dK = up(P), up(K)
if DISTINCT:
    S = np.zeros((n, 4), np.uint64); S[:, 0] = np.arange(1, n + 1, dtype=np.uint64)
    dS = up(S); dG = up(hg.g1_mul(1))
    ctx.dev("b381_g1_mul_subgroup_batch_dev", dG.data_ptr(), ctypes.c_size_t(0), dS.data_ptr(), ctypes.c_size_t(1), ctypes.c_size_t(n), dP.data_ptr())
    torch.cuda.synchronize(); del dS
dO = torch.empty(144, dtype=torch.uint8, device=dev)
res = {}; phases = {}
for spec in sys.argv[2:]:
    r, nr = map(int, spec.split(":"))
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ctx.dev("b381_g1_msm_shard_dev", dP.data_ptr(), dK.data_ptr(), ctypes.c_size_t(n), ctypes.c_int(r), ctypes.c_int(nr), dO.data_ptr())
        e1.record(st); torch.cuda.synchronize()
        t = e0.elapsed_time(e1); best = t if best is None else min(best, t)
    res[spec] = best
    ph = (ctypes.c_float * 5)()
    ctx.dev("b381_g1_msm_shard_phases_dev", dP.data_ptr(), dK.data_ptr(), ctypes.c_size_t(n), ctypes.c_int(r), ctypes.c_int(nr), dO.data_ptr(), ph)
    phases[spec] = dict(zip(["sort", "chunk_sums", "chunk_fold", "bucket_reduce", "combine"], [round(float(x), 3) for x in ph]))
print(json.dumps({"log2n": lg, "distinct_points": DISTINCT, "ms": res, "phases_ms": phases}))
