#!/usr/bin/env python3
"""A/B harness for builds of libb381.so with the same ABI: for each library given on the command line, run (in
a subprocess, B381_LIB=<lib>) the Miller-loop and final-exponentiation kernels over 2^16 resident pairs, report
the best-of-N CUDA-event times and a SHA-256 of the 2^16 x 576 B output, so that variants are compared for
speed AND bit-identity in one gpurun call.  Usage: python tools/ab_pairing.py lib1.so lib2.so ... [--n 65536]"""
import ctypes
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n, reps):
    import numpy as np
    import torch
    from bls_b200 import capi, hostgen as hg
    ctx = capi.Ctx(0)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    P = hg.g1_progression(0xB2000002, 0x9E3779B97F4A7C15, n)
    Q = hg.g2_progression(0x5EED5EED, 0xBF58476D1CE4E5B9, n)
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)
    dP, dQ = up(P), up(Q)
    dMl = torch.empty(n * 576, dtype=torch.uint8, device=dev)
    dOut = torch.empty(n * 576, dtype=torch.uint8, device=dev)
    best = [None, None]
    for r in range(reps + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        ctx.dev("b381_miller_loop_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(n), dMl.data_ptr())
        e[1].record(stream)
        ctx.dev("b381_final_exp_batch_dev", dMl.data_ptr(), ctypes.c_size_t(n), dOut.data_ptr(), None)
        e[2].record(stream)
        torch.cuda.synchronize()
        if r == 0:
            continue
        for k in range(2):
            ms = e[k].elapsed_time(e[k + 1])
            best[k] = ms if best[k] is None else min(best[k], ms)
    h = hashlib.sha256(dOut.cpu().numpy().tobytes()).hexdigest()
    hm = hashlib.sha256(dMl.cpu().numpy().tobytes()).hexdigest()
    print(json.dumps({"lib": os.path.basename(os.environ.get("B381_LIB", "libb381.so")), "n": n, "miller_ms": best[0],
                      "final_exp_ms": best[1], "pairings_per_s": n / ((best[0] + best[1]) * 1e-3),
                      "sha256_out": h[:16], "sha256_miller": hm[:16]}), flush=True)


def main():
    if os.environ.get("AB_CHILD"):
        child(int(os.environ["AB_N"]), int(os.environ["AB_REPS"]))
        return
    libs = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = 65536
    reps = 3
    for i, a in enumerate(sys.argv):
        if a == "--n": n = int(sys.argv[i + 1]); libs.remove(sys.argv[i + 1])
        if a == "--reps": reps = int(sys.argv[i + 1]); libs.remove(sys.argv[i + 1])
    for lib in libs:
        env = dict(os.environ, AB_CHILD="1", AB_N=str(n), AB_REPS=str(reps), B381_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=900)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            sys.stdout.write(json.dumps({"lib": lib, "error": r.stderr[-400:]}) + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
