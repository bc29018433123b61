#!/usr/bin/env python3
"""bench.py -- BLS12-381 pairings/s on B200 (BASELINE.json config 2: 2^16 independent pairings per GPU).

One "step" = one pass of the hot path (fused Miller loop kernel + final-exponentiation kernel)
over a batch of 2^16 synthetic (P_i, Q_i) pairs with known discrete logs.
  value : whole-job pairings/s, inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : the same steps through the public C-ABI with pinned HOST buffers: b381_pairing_batch_stream, one call for all
          steps, one batch per step (H2D of 2^16 x 304 B and D2H of 2^16 x 576 B per step inside the timed region, the
          copies of neighbouring steps overlapped with the kernels); e2e.one_call_per_step is b381_pairing_batch once per
          step (copies and kernels in series)
  roofline : integer-pipe roofline of the dominant kernel -- wide multiply-accumulates per second
          against the IMAD.WIDE.U32 issue rate measured on this GPU in the same run
          (MEASURED_PEAKS.json has no integer peak; SURVEY.md section 8d names IMAD as the bound)
  cpu_baseline : the CPU oracle (C++ restatement of the reference's pure-Go path; Go is not
          installed on this image) timed on the host cores in the same run
`--impl reference` times that CPU path alone, with all host threads, and prints the same line.

Multi-GPU (torchrun, one rank per GPU): pairings are embarrassingly parallel -- every rank runs its
own 2^16 batch, no data-path collective; scaling is "weak".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRINGS = 1 << 16
# Wide multiply-accumulates (32x32->64, IMAD.WIDE.U32) the kernels execute per unit, counted by the instrumented host build of the
# device code and pinned by tests/test_emu_logic.py::test_op_counts: an Fq product or squaring is 300 (12x12 product + 12x12
# reduction + 12 quotient words), a two-product dot product with one reduction (the rows of an Fq2 product) 444.
MACS_PER_FQ_MUL = 300
MACS_IMPL_MILLER = 2052576       # k_miller_loop, one pair (= 6 841.9 Fq multiplications of 300; 1 360 products + 3 704 dot products)
MACS_IMPL_FINAL_EXP = 1895976    # k_final_exp (= 6 319.9; 690 + 3 804), plus six integer inversions that use no multiplier
MACS_IMPL_MILLER2 = 3444480      # k_miller_loop2, two pairs sharing the accumulator
W_IMPL_MILLER = MACS_IMPL_MILLER / MACS_PER_FQ_MUL
W_IMPL_FINAL_EXP = MACS_IMPL_FINAL_EXP / MACS_PER_FQ_MUL
W_REF = 26546                    # Fq multiplications per bls.Pairing in the reference (oracle op count, tests/test_oracle_golden.py)
METRIC = "BLS12-381 pairings/sec (batches of 2^16 independent pairings per GPU)"
UNIT = "pairings/s"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_inputs(rank, n):
    from bls_b200 import hostgen as hg
    P = hg.g1_progression(0xB2000002 + 7919 * rank, 0x9E3779B97F4A7C15, n)
    Q = hg.g2_progression(0x5EED5EED + 104729 * rank, 0xBF58476D1CE4E5B9, n)
    return P, Q


def cpu_baseline(threads, target_s, P, Q):
    """time the oracle's bls.Pairing on `threads` host threads for about target_s seconds"""
    from oracle import pyoracle as orc
    m = min(P.size, 1024)
    n, t = 0, 0.0
    chunk = 16 * threads
    orc.time_pairings(P[:m], Q[:m], threads, threads)      # spin the threads up
    while t < target_s:
        t += orc.time_pairings(P[:m], Q[:m], chunk, threads)
        n += chunk
    return n / t, n, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    P, Q = make_inputs(0, 1024)
    from oracle import pyoracle as orc
    per_step = 32 * threads              # bounded sample of the 2^16-pairing step
    for _ in range(args.warmup):
        orc.time_pairings(P, Q, per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.time_pairings(P, Q, per_step, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (6x64-bit limbs)",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": "2^16 independent pairings per step; CPU arm runs a bounded sample of %d pairings "
                               "per step and reports pairings/s" % per_step, "pairings_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps x %d pairings (C++ restatement of phoreproject/bls pure-Go path; "
                                   "Go toolchain unavailable)" % (args.steps, per_step)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def pairing_source_hash():
    """SHA-256 over the sources that determine the code of k_miller_loop / k_final_exp and over the nvcc flags: the stamp that
    ties an ncu capture (profiles/traffic.json) to the library it was taken from"""
    import hashlib
    import __graft_entry__ as ge
    h = hashlib.sha256(" ".join(ge.NVCC_FLAGS).encode())
    for f in ("fp.cuh", "fp_mul_asm.inc", "constants.inc", "tower.cuh", "pairing.cuh"):
        with open(os.path.join(ROOT, "bls_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def traffic_from_profiles(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/refresh_profiles.py with the source stamp of the library it profiled).
    Returns (bytes or None, provenance): None when nothing was captured or when the sources changed since the capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None, "no capture (profiles/traffic.json missing)"
    if t.get("source_stamp") != pairing_source_hash():
        return None, "capture %s is of other sources (stamp %s, library %s): not reported" % (t.get("source"), t.get("source_stamp"), pairing_source_hash())
    return t.get(kernel), "%s, commit %s, source stamp %s" % (t.get("source"), t.get("commit"), t.get("source_stamp"))


def aggregate_extras(ctx, stream, dev, rank, world, torch, np, imad_peak):
    """aggregate-verification side of the metric (BASELINE.json configs 3-5), device-resident inputs, CUDA events:
    (a) one g1pubs VerifyAggregateCommon over 2^20 public keys (sum kernel + one 2-pair check),
    (b) a batch of 2^15 attestations x 128-key committees (Ethereum-beacon shape; 2^18 over 8 GPUs),
    (c) G1 MSMs of 2^20 and 2^22 points (BASELINE config 4) with 255-bit scalars and per-phase device times; under torchrun
        also bucket-sharded over the ranks with one all-gather of 144-byte partials (bls_b200/dist.py) and the speed-up over
        the single-rank time of the same run."""
    import torch.distributed as dist
    from bls_b200 import hostgen as hg, layout as L, dist as bd
    out = {}
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
    m = 1 << 14
    s0, d0 = 0xA66, 0x51              # same registry on every rank (replicated)
    keys = hg.g1_progression(s0, d0, m)                                   # sk_i = s0 + i d0
    sk_sum = sum(s0 + i * d0 for i in range(m)) % L.R_ORDER
    # (a) 2^20 keys = the 2^14 distinct keys tiled 64x (throughput of the reduction does not depend on the values)
    n_a = 1 << 20
    dK = up(np.resize(keys, n_a))
    dPk = torch.empty(144, dtype=torch.uint8, device=dev)
    h = 0x1234567
    Hm = hg.g2_mul(h)                                                     # stand-in for HashG2(m): the Go host hashes
    sig = hg.g2_mul(64 * sk_sum * h)
    g1one = hg.g1_mul(1)
    dOk = torch.empty(1, dtype=torch.uint8, device=dev)
    dOff = up(np.array([0, 2], np.uint32))
    dPP = torch.empty(2 * 104, dtype=torch.uint8, device=dev); dQQ = up(np.concatenate([sig, Hm]))
    hostPk = np.zeros(1, dtype=L.G1_JAC)

    def one_verify():
        ctx.dev("b381_g1_sum_dev", dK.data_ptr(), ctypes.c_size_t(n_a), dPk.data_ptr())
    t_sum = timed(one_verify)
    ctx.call("b381_d2h", ctypes.c_void_p(hostPk.ctypes.data), ctypes.c_void_p(dPk.data_ptr()), ctypes.c_size_t(144))
    agg = np.zeros(1, dtype=L.G1_AFFINE); agg["x"] = hostPk["x"]; agg["y"] = hostPk["y"]
    assert agg.tobytes() == hg.g1_mul(64 * sk_sum).tobytes(), "aggregate public key differs from the closed form"
    PP = np.concatenate([g1one, hg.g1_neg(agg)])
    dPP.copy_(up(PP))
    t_chk = timed(lambda: ctx.dev("b381_pairing_product_is_one_dev", dPP.data_ptr(), dQQ.data_ptr(), ctypes.c_size_t(2),
                                  dOff.data_ptr(), ctypes.c_size_t(1), dOk.data_ptr()))
    assert int(dOk.cpu()[0]) == 1, "aggregate verification failed on a valid signature"
    out["verify_aggregate_common_2^20_keys"] = {"sum_ms": t_sum, "pairing_check_ms": t_chk,
                                                 "verifies_per_s": 1e3 / (t_sum + t_chk),
                                                 "note": "ONE VerifyAggregateCommon at a time: 2^20-key sum + one 2-pair check (latency-bound)"}
    # the same workload as a stream: B independent aggregate verifications in flight -- B key sums back to back, then ONE launch
    # of B checks (the check's latency is paid once per B verifications)
    B = 16
    dPkB = torch.empty(B * 144, dtype=torch.uint8, device=dev)
    dOkB = torch.empty(B, dtype=torch.uint8, device=dev)
    dOffB = up((2 * np.arange(B + 1)).astype(np.uint32))
    dPPB = up(np.tile(PP, B)); dQQB = up(np.tile(np.concatenate([sig, Hm]), B))

    def stream_verifies():
        for b in range(B):
            ctx.dev("b381_g1_sum_dev", dK.data_ptr(), ctypes.c_size_t(n_a), dPkB.data_ptr() + 144 * b)
        ctx.dev("b381_pairing_product_is_one_dev", dPPB.data_ptr(), dQQB.data_ptr(), ctypes.c_size_t(2 * B), dOffB.data_ptr(),
                ctypes.c_size_t(B), dOkB.data_ptr())
    t_stream = timed(stream_verifies)
    assert dOkB.cpu().numpy().all(), "a streamed aggregate verification failed on a valid signature"
    out["verify_aggregate_common_2^20_keys_stream"] = {"in_flight": B, "ms_per_batch": t_stream, "verifies_per_s": B * 1e3 / t_stream,
                                                        "note": "%d aggregate verifications of 2^20 keys each per batch: key sums + one launch of %d checks "
                                                                "(the pairs handed to the check are the precomputed (G1, sig), (-pk, H) of the valid case)" % (B, B)}
    # (b) attestation batch: 64 distinct (committee, message) templates tiled to 2^14 attestations, 1 in 64 corrupted
    natt, comm, ntmpl = 1 << 15, 128, 64        # BASELINE config 5: 2^18 attestations over 8 GPUs = 2^15 per GPU
    rng = np.random.RandomState(1 + rank)
    tk = rng.randint(0, m, size=(ntmpl, comm))
    hs = [0x77 + 5 * j for j in range(8)]
    Hs = np.concatenate([hg.g2_mul(x) for x in hs])
    sigs, expect = [], []
    for t in range(ntmpl):
        sk = sum(s0 + int(i) * d0 for i in tk[t]) % L.R_ORDER
        bad = t == 7
        sigs.append(hg.g2_mul((sk + (1 if bad else 0)) * hs[t % 8])); expect.append(0 if bad else 1)
    kidx = np.tile(tk.reshape(-1), natt // ntmpl).astype(np.uint32)
    koff = (np.arange(natt + 1) * comm).astype(np.uint32)
    dA = [up(x) for x in (keys, kidx, koff, np.tile(np.concatenate(sigs), natt // ntmpl), Hs,
                          np.tile(np.arange(ntmpl) % 8, natt // ntmpl).astype(np.uint32))]
    dOkA = torch.empty(natt, dtype=torch.uint8, device=dev)
    t_att = timed(lambda: ctx.dev("b381_verify_aggregate_common_batch_dev", *[x.data_ptr() for x in dA], ctypes.c_size_t(natt),
                                  ctypes.c_size_t(keys.size), ctypes.c_size_t(Hs.size), dOkA.data_ptr()), reps=2)
    assert dOkA.cpu().numpy()[:ntmpl].tolist() == expect, "attestation batch verdicts differ from construction"
    ta = torch.tensor([t_att], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    w_att = comm * 10 + 600 + (MACS_IMPL_MILLER2 + MACS_IMPL_FINAL_EXP) / MACS_PER_FQ_MUL   # key aggregation + normalisation + shared-accumulator Miller loop + final exp
    out["attestation_batch_2^15_x_128_keys"] = {
        "ms": float(ta.item()), "aggregate_verifies_per_s": world * natt / (float(ta.item()) * 1e-3),
        "note": "BASELINE config 5 = 2^18 attestations over 8 GPUs = 2^15 per GPU; pre-hashed message points, resident key registry",
        "roofline": {"kernels": "k_attest_pairs + k_miller_loop2 + k_final_exp_is_one", "fq_mul_per_attestation": w_att,
                     "achieved": natt * w_att * MACS_PER_FQ_MUL / (t_att * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                     "unit": "T wide-MAC/s", "frac": natt * w_att * MACS_PER_FQ_MUL / (t_att * 1e-3) / imad_peak}}
    # (b'') the same attestation batch as ONE boolean: random linear combination grouped by message (nmsg + 1 Miller loops and one
    # final exponentiation per batch; what scales with the batch is the committee sums, one 64-bit scalar multiplication per
    # attestation, the group-by-message sum and a 64-bit-weight G2 MSM).  Timed on the all-valid batch (must be accepted); the
    # batch with the corrupted template must be rejected.
    sig_ok = []
    for t in range(ntmpl):
        sk = sum(s0 + int(i) * d0 for i in tk[t]) % L.R_ORDER
        sig_ok.append(hg.g2_mul(sk * hs[t % 8]))
    dSigOk = up(np.tile(np.concatenate(sig_ok), natt // ntmpl))
    wts = np.zeros((natt, 4), np.uint64)
    wts[:, 0] = np.random.Generator(np.random.PCG64(7 + rank)).integers(1, 1 << 64, size=natt, dtype=np.uint64)
    dW = up(wts)
    dOk1 = torch.zeros(1, dtype=torch.uint8, device=dev)
    ctx.call("b381_set_rlc_weight_bits", ctypes.c_int(64))

    def att_rlc(sig_buf):
        ctx.dev("b381_verify_aggregate_common_rlc_dev", dA[0].data_ptr(), dA[1].data_ptr(), dA[2].data_ptr(), sig_buf.data_ptr(), dA[4].data_ptr(),
                dA[5].data_ptr(), dW.data_ptr(), ctypes.c_size_t(natt), ctypes.c_size_t(keys.size), ctypes.c_size_t(Hs.size), dOk1.data_ptr())
    att_rlc(dA[3]); torch.cuda.synchronize()
    assert int(dOk1.item()) == 0, "the batch with corrupted attestations must be rejected"
    t_rlc = timed(lambda: att_rlc(dSigOk), reps=3)
    assert int(dOk1.item()) == 1, "the all-valid attestation batch must be accepted"
    tr = torch.tensor([t_rlc], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    out["attestation_batch_2^15_x_128_keys_one_boolean"] = {
        "ms": float(tr.item()), "aggregate_verifies_per_s": world * natt / (float(tr.item()) * 1e-3), "distinct_messages": int(Hs.size),
        "weight_bits": 64, "note": "b381_verify_aggregate_common_rlc_dev: batch verification by a random linear combination grouped by message "
                                   "(soundness error 2^-64 per batch); on rejection the per-attestation call above locates the offenders"}
    ctx.call("b381_set_rlc_weight_bits", ctypes.c_int(255))
    del dSigOk, dW
    # (b') g1pubs.VerifyWithDomain from wire bytes: 2^16 (public key, message hash, signature) triples per GPU; deserialisation,
    # subgroup checks, HashG2WithDomain and the 2-pair check all on the device.  Inputs are made with the engine itself
    # (PrivToPub / SignWithDomain / Serialize batches, each parity-tested against the oracle); one triple in 64 is corrupted.
    nw = 1 << 16
    one, zero, NW = ctypes.c_size_t(1), ctypes.c_size_t(0), ctypes.c_size_t(nw)
    Kw, _ = hg.splitmix_scalars(7 + rank, 1 << 12)
    dKw = up(np.resize(Kw, (nw, 4)))
    rngw = np.random.RandomState(100 + rank)
    msgs_w = rngw.randint(0, 256, (nw, 32), dtype=np.uint8)
    dM = up(msgs_w); dDom = up(np.arange(8, dtype=np.uint8)); dG = up(hg.g1_mul(1))
    dPub = torch.empty(nw * 104, dtype=torch.uint8, device=dev); dPubC = torch.empty(nw * 48, dtype=torch.uint8, device=dev)
    dH = torch.empty(nw * 200, dtype=torch.uint8, device=dev); dSg = torch.empty(nw * 200, dtype=torch.uint8, device=dev)
    dSgC = torch.empty(nw * 96, dtype=torch.uint8, device=dev); dOkW = torch.empty(nw, dtype=torch.uint8, device=dev)
    ctx.dev("b381_g1_mul_batch_dev", dG.data_ptr(), zero, dKw.data_ptr(), one, NW, dPub.data_ptr())
    ctx.dev("b381_g1_compress_batch_dev", dPub.data_ptr(), NW, dPubC.data_ptr())
    ctx.dev("b381_hash_g2_with_domain_batch_dev", dM.data_ptr(), dDom.data_ptr(), zero, NW, dH.data_ptr())
    ctx.dev("b381_g2_mul_batch_dev", dH.data_ptr(), one, dKw.data_ptr(), one, NW, dSg.data_ptr())
    ctx.dev("b381_g2_compress_batch_dev", dSg.data_ptr(), NW, dSgC.data_ptr())
    dM[63 * 32::64 * 32] ^= 1                                     # flip a bit of every 64th message
    t_wire = timed(lambda: ctx.dev("b381_verify_with_domain_batch_dev", dPubC.data_ptr(), dM.data_ptr(), dDom.data_ptr(), zero,
                                   dSgC.data_ptr(), NW, dOkW.data_ptr()), reps=2)
    okw = dOkW.cpu().numpy()
    assert okw.reshape(-1, 64)[:, :63].all() and not okw.reshape(-1, 64)[:, 63].any(), "wire-level verdicts differ from construction"
    tw = torch.tensor([t_wire], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    out["verify_with_domain_wire_2^16"] = {"ms": float(tw.item()), "verifies_per_s": world * nw / (float(tw.item()) * 1e-3),
                                           "bytes_in_per_verify": 48 + 32 + 96}
    del dKw, dM, dPub, dPubC, dH, dSg, dSgC, dOkW
    # (c) BASELINE config 4: 2^22-point G1 MSM, 255-bit scalars (and the 2^20 size of config 3).  Closed-form check on the tiled
    # points: sum_i k_(i mod 4096) sk_(i mod m).  Single GPU: all windows on this rank, with the device time of each phase.
    # Under torchrun the bucket space is sharded over the ranks (rank g owns the windows w = g mod G), one all-gather of
    # 144-byte partials + a fold on every rank; the speed-up is against the single-rank time measured in the same run.
    hostPk2 = np.zeros(1, dtype=L.G1_JAC)
    PH = ("sort", "chunk_sums", "chunk_fold", "bucket_reduce", "combine")
    for lg in (20, 22):
        n_c = 1 << lg
        dKp = dK if lg == 20 else up(np.resize(keys, n_c))
        # one independent uniform scalar < 2^254 per point (tiling a few thousand scalars would put ~1000 points into each of a few
        # buckets, a skewed case the engine handles through its tree rounds but not what a weighted aggregation looks like)
        K = np.random.Generator(np.random.PCG64(99 + lg)).integers(0, 1 << 64, size=(n_c, 4), dtype=np.uint64)
        K[:, 3] &= np.uint64((1 << 62) - 1)
        dKs = up(K)
        dOut = torch.empty(144, dtype=torch.uint8, device=dev)
        # closed form on the tiled points P_i = (s0 + (i mod m) d0) G: sum_j (s0 + j d0) * (sum of the scalars of residue j), the
        # residue sums taken exactly on 32-bit halves
        Kr = K.reshape(n_c // m, m, 4)
        lo = (Kr & np.uint64(0xFFFFFFFF)).sum(axis=0, dtype=np.uint64); hi = (Kr >> np.uint64(32)).sum(axis=0, dtype=np.uint64)
        S = 0
        for j in range(m):
            kj = sum((int(lo[j, w]) + (int(hi[j, w]) << 32)) << (64 * w) for w in range(4))
            S += kj * (s0 + j * d0)
        want = hg.g1_mul(S % L.R_ORDER).tobytes()
        del K, Kr

        def check(what):
            ctx.call("b381_d2h", ctypes.c_void_p(hostPk2.ctypes.data), ctypes.c_void_p(dOut.data_ptr()), ctypes.c_size_t(144))
            agg["x"] = hostPk2["x"]; agg["y"] = hostPk2["y"]
            assert agg.tobytes() == want, what + " differs from the closed form"
        t_msm = timed(lambda: ctx.dev("b381_g1_msm_dev", dKp.data_ptr(), dKs.data_ptr(), ctypes.c_size_t(n_c), dOut.data_ptr()))
        check("MSM result")
        ph = (ctypes.c_float * 5)()
        ctx.dev("b381_g1_msm_shard_phases_dev", dKp.data_ptr(), dKs.data_ptr(), ctypes.c_size_t(n_c), 0, 1, dOut.data_ptr(), ph)
        c_bits = max(4, min(16, lg - 5)); W = (255 + c_bits) // c_bits          # signed digits: one more bit for the last carry
        macs = n_c * W * (8 * MACS_PER_FQ_MUL + 2 * 228)  # XYZZ mixed addition per point and window: 8 products (300 MACs) + 2 squarings (228, fp_sqr_v)
        blk = {"n": n_c, "scalar_bits": 255, "window_bits": c_bits, "windows": W, "ms": t_msm, "points_per_s": n_c / (t_msm * 1e-3),
               "phases_ms": {k: float(v) for k, v in zip(PH, ph)},
               "roofline": {"kernel": "k_msm_chunk_sum", "bound": "int32-imad", "fq_mul_per_point": W * 10, "wide_macs_per_point": W * (8 * MACS_PER_FQ_MUL + 2 * 228),
                            "achieved": macs / (ph[1] * 1e-3) / 1e12, "peak": imad_peak / 1e12, "unit": "T wide-MAC/s",
                            "frac": macs / (ph[1] * 1e-3) / imad_peak,
                            "hbm_gbs_sanity": n_c * W * (4 + 104) / (ph[1] * 1e-3) / 1e9}}
        if world > 1:
            dParts = torch.empty(world * 144, dtype=torch.uint8, device=dev)
            t_sh = timed(lambda: bd.msm_bucket_sharded_dev(ctx, dKp, dKs, n_c, dParts, dOut))
            ts = torch.tensor([t_sh], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            check("bucket-sharded MSM")
            ph2 = (ctypes.c_float * 5)()
            ctx.dev("b381_g1_msm_shard_phases_dev", dKp.data_ptr(), dKs.data_ptr(), ctypes.c_size_t(n_c), rank, world, dOut.data_ptr(), ph2)
            blk["bucket_sharded"] = {"ranks": world, "ms": float(ts.item()), "speedup_vs_one_gpu": t_msm / float(ts.item()),
                                     "exchange": "one all_gather of %d x 144 B (NCCL) + fold on every rank" % world,
                                     "phases_ms_this_rank": {k: float(v) for k, v in zip(PH, ph2)}}
        out["g1_msm_2^%d_255bit" % lg] = blk
        del dKs
        if lg == 22:
            del dKp
    # BASELINE config 3 as one number: AggregateVerify of 2^20 signatures = 2^20-point MSM (random weights) + one 2-pair check
    t3 = out["g1_msm_2^20_255bit"]["ms"] + t_chk
    out["aggregate_verify_2^20_weighted"] = {"msm_ms": out["g1_msm_2^20_255bit"]["ms"], "pairing_check_ms": t_chk, "verifies_per_s": 1e3 / t3}
    return out


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_engine(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bls_b200 import capi, layout as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = capi.Ctx(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    n = N_PAIRINGS
    P, Q = make_inputs(rank, n)
    # pinned host staging (e2e) and resident device copies (value)
    hP = torch.empty(P.nbytes, dtype=torch.uint8).pin_memory(); hP.numpy()[:] = P.view(np.uint8).reshape(-1)
    hQ = torch.empty(Q.nbytes, dtype=torch.uint8).pin_memory(); hQ.numpy()[:] = Q.view(np.uint8).reshape(-1)
    hOut = torch.empty(n * 576, dtype=torch.uint8).pin_memory()
    dP = hP.to(dev); dQ = hQ.to(dev)
    dOut = torch.empty(n * 576, dtype=torch.uint8, device=dev)
    dMl = torch.empty(n * 576, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        ctx.dev("b381_pairing_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(n), dOut.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- value: K timed steps, L2 flushed between them, device time from CUDA events -------------
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- per-kernel durations for the roofline (events on the launching stream) ------------------
    kt = {"k_miller_loop": 0.0, "k_final_exp": 0.0}
    reps = max(3, min(args.steps, 5))
    for _ in range(reps):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        flush.fill_(1)
        e0.record(stream)
        ctx.dev("b381_miller_loop_batch_dev", dP.data_ptr(), dQ.data_ptr(), ctypes.c_size_t(n), dMl.data_ptr())
        e1.record(stream)
        ctx.dev("b381_final_exp_batch_dev", dMl.data_ptr(), ctypes.c_size_t(n), dOut.data_ptr(), None)
        e2.record(stream)
        torch.cuda.synchronize()
        kt["k_miller_loop"] += e0.elapsed_time(e1) / reps
        kt["k_final_exp"] += e1.elapsed_time(e2) / reps
    # integer-pipe peak: IMAD.WIDE.U32 issue rate, measured now on this GPU at its current clocks
    pb, pt, pi = 148 * 8, 256, 8192
    probe_out = torch.empty(pb * pt, dtype=torch.int32, device=dev)
    best = None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.dev("b381_imad_probe_dev", probe_out.data_ptr(), pb, pt, pi)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    imad_peak_plain = pb * pt * pi * 8 / (best * 1e-3)    # wide MACs per second, independent 64-bit accumulates
    # the engine's own multiplier in a register-only dependent chain (carry-chained IMAD.WIDE.U32.X): the practical peak
    fb, ft, fi = 148 * 4, 256, 2048
    probe2 = torch.empty(fb * ft, dtype=torch.int32, device=dev)
    best2 = None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.dev("b381_fpmul_probe_dev", probe2.data_ptr(), fb, ft, fi)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best2 = ms if best2 is None else min(best2, ms)
    imad_peak_chain = fb * ft * fi * MACS_PER_FQ_MUL / (best2 * 1e-3)
    imad_peak = max(imad_peak_plain, imad_peak_chain)

    # ---- e2e: public host-buffer call, H2D + kernels + D2H inside the timed region ----------------
    pP = ctypes.c_void_p(hP.data_ptr()); pQ = ctypes.c_void_p(hQ.data_ptr()); pO = ctypes.c_void_p(hOut.data_ptr())

    def e2e_step():
        ctx.call("b381_pairing_batch", pP, pQ, ctypes.c_size_t(n), pO)

    e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_single = world * n * e2e_steps / float(te.item())
    # the same steps as ONE call of the streaming form: every step's inputs go up from pinned host memory and its results come
    # down inside the timed region, the copies of neighbouring steps overlapped with the kernels (b381_pairing_batch_stream)
    hPs = torch.empty(e2e_steps * P.nbytes, dtype=torch.uint8).pin_memory(); hQs = torch.empty(e2e_steps * Q.nbytes, dtype=torch.uint8).pin_memory()
    hOs = torch.empty(e2e_steps * n * 576, dtype=torch.uint8).pin_memory()
    for k in range(e2e_steps):
        hPs.numpy()[k * P.nbytes:(k + 1) * P.nbytes] = hP.numpy(); hQs.numpy()[k * Q.nbytes:(k + 1) * Q.nbytes] = hQ.numpy()

    def e2e_stream():
        ctx.call("b381_pairing_batch_stream", ctypes.c_void_p(hPs.data_ptr()), ctypes.c_void_p(hQs.data_ptr()), ctypes.c_size_t(e2e_steps * n),
                 ctypes.c_size_t(n), ctypes.c_void_p(hOs.data_ptr()))
    e2e_stream()
    barrier()
    t0 = time.perf_counter()
    e2e_stream()
    barrier()
    t_e2s = time.perf_counter() - t0
    assert hOs.numpy()[-n * 576:].tobytes() == hOut.numpy().tobytes(), "streamed results differ from the single-batch call"
    te = torch.tensor([t_e2s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(te.item())
    del hPs, hQs, hOs

    extras = aggregate_extras(ctx, stream, dev, rank, world, torch, np, imad_peak) if not args.no_aggregate else None
    # ---- result check (sample vs the oracle) + CPU baseline on rank 0 ---------------------------
    if rank == 0:
        from oracle import pyoracle as orc
        got = np.frombuffer(hOut.numpy().tobytes(), dtype=np.uint64).reshape(n, 2, 3, 2, 6)
        idx = np.arange(0, n, n // 16)
        exp = orc.pairing_batch(P[idx], Q[idx], threads=host_threads())
        assert got[idx].tobytes() == exp.tobytes(), "engine output differs from the oracle"
        dgot = np.frombuffer(dOut.cpu().numpy().tobytes(), dtype=np.uint64).reshape(n, 2, 3, 2, 6)
        assert dgot[idx].tobytes() == exp.tobytes(), "resident-path output differs from the oracle"

        dom = max(kt, key=kt.get)
        w_dom = W_IMPL_FINAL_EXP if dom == "k_final_exp" else W_IMPL_MILLER
        achieved = n * w_dom * MACS_PER_FQ_MUL / (kt[dom] * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32x12 limbs (384-bit Montgomery integers)",
            "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: batch of 2^16 independent BLS12-381 pairings per GPU "
                                   "(P_i=(s+i*d)G1, Q_i=(s'+i*d')G2), affine inputs, Fq12 outputs",
                       "pairings_per_gpu_per_step": n, "parallelism": "batch-sharded x%d, no collective" % world,
                       "l2": "256 MiB buffer written between timed steps (L2 flush); path is compute-bound",
                       "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (104 + 200),
                    "d2h_bytes_per_step": n * 576, "steps": e2e_steps,
                    "api": "b381_pairing_batch_stream (host pointers, pinned; one call for all steps, batch = one step, copies of "
                           "neighbouring steps overlapped with the kernels)",
                    "one_call_per_step": {"value": e2e_single, "api": "b381_pairing_batch (host pointers, pinned; H2D, kernels, D2H in series)"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "int32-imad (no HBM or tensor bound: ~4.4M wide MACs per 880 B of I/O)",
                "kernel": dom, "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
                "unit": "T wide-MAC/s (IMAD.WIDE.U32)", "frac": achieved / imad_peak,
                "peak_source": "measured in this run on this GPU: max of k_fpmul_probe (dependent Fq-multiplication chain, "
                               "%.2f T/s) and k_imad_probe (independent mad.wide chains, %.2f T/s); MEASURED_PEAKS.json has no "
                               "integer peak" % (imad_peak_chain / 1e12, imad_peak_plain / 1e12),
                "traffic": traffic_from_profiles(dom)[0], "traffic_source": traffic_from_profiles(dom)[1],
                "kernels_ms": kt,
                "fq_mul_per_pairing": {"impl_miller": W_IMPL_MILLER, "impl_final_exp": W_IMPL_FINAL_EXP, "reference": W_REF},
                "ref_equivalent_frac": (n * W_REF * MACS_PER_FQ_MUL / ((kt["k_miller_loop"] + kt["k_final_exp"]) * 1e-3)) / imad_peak,
                "hbm_gbs_sanity": n * (104 + 200 + 576 * 3) / ((kt["k_miller_loop"] + kt["k_final_exp"]) * 1e-3) / 1e9,
            },
        }
        line["aggregate"] = extras
        threads = host_threads()
        cv, cn, ct = cpu_baseline(threads, args.cpu_seconds, P, Q)
        line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d pairings in %.1f s on %d host threads (C++ restatement of "
                                          "phoreproject/bls pure-Go path; Go toolchain unavailable)" % (cn, ct, threads)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract goes to the real stdout; everything else a library prints on fd 1
    (e.g. NCCL's version banner) was redirected to stderr in main()"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the cpu_baseline sample")
    ap.add_argument("--no-aggregate", action="store_true", help="skip the aggregate-verification extras (configs 3-5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
