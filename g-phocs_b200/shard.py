"""Locus sharding across ranks and the per-step all-reduce payload.

Loci are conditionally independent given the model parameters, so rank r owns the contiguous block
shard_range(L, r, world) — the same split OpenMP's schedule(static) makes in the reference
(/root/reference/src/MultiCoreUtils.h:8) — and the only cross-rank traffic per global proposal is one sum of
    [ sum data lnL | sum genealogy lnL, coal_stats[Q], num_coals[Q], mig_stats[B], num_migs[B] ]
(computeTotalStats, patch.c:2134-2164; the atomics of GPhoCS.c:3807-3836).  Integer counts travel as
doubles, exact below 2^53.  Host-side plumbing only: no likelihood arithmetic here.
"""
import numpy as np


def shard_range(num_loci, rank, world):
    """[lo, hi) of rank's block: ceil(L/world) loci per rank, the last blocks may be shorter (static schedule)."""
    chunk = -(-num_loci // world)
    lo = min(rank * chunk, num_loci)
    return lo, min(lo + chunk, num_loci)


def payload_len(Q, B):
    return 2 + 2 * Q + 2 * B


def pack_payload(sum_data_lnl, sum_gen_lnl, total_coal, total_num_coals, total_mig, total_num_migs):
    Q, B = len(total_coal), len(total_mig)
    v = np.zeros(payload_len(Q, B))
    v[0], v[1] = sum_data_lnl, sum_gen_lnl
    v[2:2 + Q] = total_coal
    v[2 + Q:2 + 2 * Q] = total_num_coals
    v[2 + 2 * Q:2 + 2 * Q + B] = total_mig
    v[2 + 2 * Q + B:] = total_num_migs
    return v


def unpack_payload(v, Q, B):
    v = np.asarray(v, np.float64)
    return dict(sum_data_lnl=float(v[0]), sum_gen_lnl=float(v[1]), total_coal=v[2:2 + Q].copy(),
                total_num_coals=np.rint(v[2 + Q:2 + 2 * Q]).astype(np.int64),
                total_mig=v[2 + 2 * Q:2 + 2 * Q + B].copy(),
                total_num_migs=np.rint(v[2 + 2 * Q + B:2 + 2 * Q + 2 * B]).astype(np.int64))


def all_reduce_payload(tensor):
    """In-place sum over ranks of a torch tensor holding the payload (NCCL on GPU tensors, gloo on CPU ones)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
