"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the path shards.

* independent pairings / attestations: contiguous tiles per rank, no data-path collective
  (`tile`), results optionally gathered (`gather_bytes`);
* one large MSM: bucket-sharded -- rank g owns the windows {w : w mod G == g} of the Pippenger
  bucket space, emits one Jacobian partial (144 B), and the single exchange step is an all-gather
  of the G partials followed by a local fold on every rank.  NCCL has no user-defined reduction,
  so the north star's "all-reduce of partial bucket sums" is all-gather + fold.

* random-linear-combination batch verification: every rank forms the Miller product of its own triples
  (576 bytes, no final exponentiation), one all-gather, and every rank finishes with the product of the partials and
  a single final exponentiation (`verify_rlc_sharded`).

The arithmetic is always the engine's (a `capi.Ctx`, or in the CPU test-suite a stand-in with the
same two methods); this module only moves 144-byte partials.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import layout as L


def tile(n, rank, world):
    """contiguous [lo, hi) of n independent units for this rank (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_bytes(local, group=None):
    """all-gather equally sized uint8 tensors (per-rank validity bitmaps, 576-byte results, ...)"""
    world = dist.get_world_size(group)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local, group=group)
    return torch.cat(out)


def all_ok(flag_tensor, group=None):
    """min-reduce of a per-rank 'every check passed' flag"""
    t = flag_tensor.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def msm_bucket_sharded(engine, points, scalars, group=None):
    """host-buffer form: numpy points/scalars (replicated on every rank) -> normalised G1_JAC sum.
    `engine` provides g1_msm_shard(p, k, rank, nranks) and g1_fold(parts) (capi.Ctx does)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    part = engine.g1_msm_shard(points, scalars, rank, world)
    mine = torch.from_numpy(np.ascontiguousarray(part).view(np.uint8).reshape(-1).copy())
    parts = gather_bytes(mine, group).numpy().view(L.G1_JAC)
    return engine.g1_fold(parts)


def msm_bucket_sharded_dev(ctx, d_points, d_scalars, n, d_parts, d_out, group=None):
    """device-resident form used by bench.py under NCCL: d_* are CUDA uint8 tensors; d_parts holds
    world x 144 B; everything is enqueued on the ctx stream (= torch's current stream)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = d_parts[rank * 144:(rank + 1) * 144]
    ctx.dev("b381_g1_msm_shard_dev", d_points.data_ptr(), d_scalars.data_ptr(), ctypes.c_size_t(n), rank, world,
            mine.data_ptr())
    dist.all_gather_into_tensor(d_parts, mine.clone(), group=group)
    ctx.dev("b381_g1_fold_dev", d_parts.data_ptr(), ctypes.c_size_t(world), d_out.data_ptr())
    return d_out


def verify_rlc_sharded(engine, pubs, msg_points, sigs, weights, group=None):
    """One boolean for the union of every rank's (public key, message point, signature) triples.  Each rank passes ITS
    tile; `engine` provides verify_rlc_partial(pub, h, sig, w) -> (one Fq12 value as 72 u64, valid) and
    fp12_product_final_exp_is_one(parts) (capi.Ctx does).  The single exchange step is an all-gather of 576 + 1 bytes."""
    part, valid = engine.verify_rlc_partial(pubs, msg_points, sigs, weights)
    mine = torch.from_numpy(np.concatenate([np.ascontiguousarray(part).view(np.uint8).reshape(-1), np.array([valid], np.uint8)]).copy())
    allp = gather_bytes(mine, group).numpy().reshape(-1, 577)
    if not allp[:, 576].all():
        return False
    return bool(engine.fp12_product_final_exp_is_one(np.ascontiguousarray(allp[:, :576]).view(np.uint64).reshape(-1, 72)))
