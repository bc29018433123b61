#!/usr/bin/env python3
"""Device-resident timings (CUDA events, best of 3) of the wire-format / scalar-multiplication kernels:
python tools/codec_bench.py [--n 65536] -> one JSON object on stdout."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bls_b200 import capi, hostgen as hg, layout as L

n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 65536
ctx = capi.Ctx(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


m = min(n, 4096)
P1 = np.resize(hg.g1_progression(3, 5, m), n); P2 = np.resize(hg.g2_progression(7, 11, m), n)
K, _ = hg.splitmix_scalars(5, m)
K = np.resize(K, (n, 4))
dP1, dP2, dK = up(P1), up(P2), up(K)
dC1 = torch.empty(n * 48, dtype=torch.uint8, device=dev); dC2 = torch.empty(n * 96, dtype=torch.uint8, device=dev)
dO1 = torch.empty(n * 104, dtype=torch.uint8, device=dev); dO2 = torch.empty(n * 200, dtype=torch.uint8, device=dev)
dS = torch.empty(n, dtype=torch.uint8, device=dev)
N = ctypes.c_size_t(n)
one = ctypes.c_size_t(1)
res = {"n": n}
res["g1_compress_ms"] = timed(lambda: ctx.dev("b381_g1_compress_batch_dev", dP1.data_ptr(), N, dC1.data_ptr()))
res["g2_compress_ms"] = timed(lambda: ctx.dev("b381_g2_compress_batch_dev", dP2.data_ptr(), N, dC2.data_ptr()))
for chk in (0, 1):
    tag = "checked" if chk else "unchecked"
    res["g1_decompress_%s_ms" % tag] = timed(lambda: ctx.dev("b381_g1_decompress_batch_dev", dC1.data_ptr(), N, chk, dO1.data_ptr(), dS.data_ptr()))
    assert not dS.any().item() and bytes(dO1.cpu().numpy()) == P1.tobytes()
    res["g2_decompress_%s_ms" % tag] = timed(lambda: ctx.dev("b381_g2_decompress_batch_dev", dC2.data_ptr(), N, chk, dO2.data_ptr(), dS.data_ptr()))
    assert not dS.any().item() and bytes(dO2.cpu().numpy()) == P2.tobytes()
res["g1_mul_ms"] = timed(lambda: ctx.dev("b381_g1_mul_batch_dev", dP1.data_ptr(), one, dK.data_ptr(), one, N, dO1.data_ptr()))
res["g2_mul_ms"] = timed(lambda: ctx.dev("b381_g2_mul_batch_dev", dP2.data_ptr(), one, dK.data_ptr(), one, N, dO2.data_ptr()))
dO1b = torch.empty_like(dO1); dO2b = torch.empty_like(dO2)
ctx.dev("b381_g1_mul_batch_dev", dP1.data_ptr(), one, dK.data_ptr(), one, N, dO1.data_ptr())
res["g1_mul_subgroup_ms"] = timed(lambda: ctx.dev("b381_g1_mul_subgroup_batch_dev", dP1.data_ptr(), one, dK.data_ptr(), one, N, dO1b.data_ptr()))
res["g2_mul_subgroup_ms"] = timed(lambda: ctx.dev("b381_g2_mul_subgroup_batch_dev", dP2.data_ptr(), one, dK.data_ptr(), one, N, dO2b.data_ptr()))
assert torch.equal(dO1, dO1b) and torch.equal(dO2, dO2b), "endomorphism ladder differs from the plain ladder"
rng = np.random.RandomState(1)
dM = up(rng.randint(0, 256, (n, 32), dtype=np.uint8)); dD = up(np.arange(8, dtype=np.uint8))
res["hash_g2_with_domain_ms"] = timed(lambda: ctx.dev("b381_hash_g2_with_domain_batch_dev", dM.data_ptr(), dD.data_ptr(), ctypes.c_size_t(0), N, dO2.data_ptr()))
# wire-level VerifyWithDomain: valid triples made with the engine itself
sk = dK
dPub = torch.empty(n * 104, dtype=torch.uint8, device=dev); dPubC = torch.empty(n * 48, dtype=torch.uint8, device=dev)
G = up(hg.g1_mul(1))
ctx.dev("b381_g1_mul_batch_dev", G.data_ptr(), ctypes.c_size_t(0), dK.data_ptr(), one, N, dPub.data_ptr())
ctx.dev("b381_g1_compress_batch_dev", dPub.data_ptr(), N, dPubC.data_ptr())
dH = torch.empty(n * 200, dtype=torch.uint8, device=dev); dSig = torch.empty(n * 200, dtype=torch.uint8, device=dev)
dSigC = torch.empty(n * 96, dtype=torch.uint8, device=dev); dOk = torch.empty(n, dtype=torch.uint8, device=dev)
ctx.dev("b381_hash_g2_with_domain_batch_dev", dM.data_ptr(), dD.data_ptr(), ctypes.c_size_t(0), N, dH.data_ptr())
ctx.dev("b381_g2_mul_batch_dev", dH.data_ptr(), one, dK.data_ptr(), one, N, dSig.data_ptr())
ctx.dev("b381_g2_compress_batch_dev", dSig.data_ptr(), N, dSigC.data_ptr())
res["verify_with_domain_wire_ms"] = timed(lambda: ctx.dev("b381_verify_with_domain_batch_dev", dPubC.data_ptr(), dM.data_ptr(), dD.data_ptr(),
                                                             ctypes.c_size_t(0), dSigC.data_ptr(), N, dOk.data_ptr()), reps=2)
assert bool(dOk.all().item()), "valid signatures must verify"
# random-linear-combination variants: one boolean per batch
W = np.zeros((n, 4), np.uint64); W[:, 0] = rng.randint(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
dW = up(W); dOk1 = torch.zeros(8, dtype=torch.uint8, device=dev)
res["verify_with_domain_wire_rlc_ms"] = timed(lambda: ctx.dev("b381_verify_with_domain_rlc_batch_dev", dPubC.data_ptr(), dM.data_ptr(), dD.data_ptr(),
                                                                 ctypes.c_size_t(0), dSigC.data_ptr(), dW.data_ptr(), N, dOk1.data_ptr()), reps=2)
assert int(dOk1[0].item()) == 1, "valid batch must pass the RLC check"
res["verify_rlc_resident_points_ms"] = timed(lambda: ctx.dev("b381_verify_rlc_dev", dPub.data_ptr(), dH.data_ptr(), dSig.data_ptr(), dW.data_ptr(), N,
                                                                dOk1.data_ptr()), reps=2)
assert int(dOk1[0].item()) == 1
# SWU hashing and the plain Verify from wire bytes (64-byte messages)
ml = 64
dMsg = up(rng.randint(0, 256, (n, ml), dtype=np.uint8)); dOff = up((np.arange(n + 1, dtype=np.uint64) * ml))
res["hash_g1_ms"] = timed(lambda: ctx.dev("b381_hash_g1_batch_dev", dMsg.data_ptr(), dOff.data_ptr(), N, dO1.data_ptr()))
res["hash_g2_ms"] = timed(lambda: ctx.dev("b381_hash_g2_batch_dev", dMsg.data_ptr(), dOff.data_ptr(), N, dH.data_ptr()))
ctx.dev("b381_g2_mul_batch_dev", dH.data_ptr(), one, dK.data_ptr(), one, N, dSig.data_ptr())
ctx.dev("b381_g2_compress_batch_dev", dSig.data_ptr(), N, dSigC.data_ptr())
res["g1pubs_verify_wire_ms"] = timed(lambda: ctx.dev("b381_g1pubs_verify_batch_dev", dPubC.data_ptr(), dMsg.data_ptr(), dOff.data_ptr(),
                                                        dSigC.data_ptr(), N, dOk.data_ptr()), reps=2)
assert bool(dOk.all().item()), "valid signatures must verify (g1pubs.Verify)"
for k in list(res):
    if k.endswith("_ms"):
        res[k.replace("_ms", "_per_s")] = n / (res[k] * 1e-3)
print(json.dumps(res))
