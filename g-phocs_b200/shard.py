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


class PipelinedAllReduce:
    """The per-step payload all-reduce taken off the evaluation stream.

    The payload of step i is summed over ranks on a side stream while the kernels of step i+1 run: the evaluation stream
    only records an event after it has written the payload and, `depth` steps later, waits for the event that marks
    that buffer's all-reduce as finished (long since, in practice).  Results are therefore available one step late —
    fine for what consumes them (running totals for the trace line, global proposals that are decided once per sweep).
    CPU tensors (gloo) use asynchronous work handles instead of streams; same interface."""

    def __init__(self, length, device, depth=2, dtype=None, group=None):
        import torch
        self.torch = torch
        self.group = group   # process group of the all-reduce (None: the default group)
        self.cuda = torch.device(device).type == "cuda"
        self.buffers = [torch.zeros(length, dtype=dtype or torch.float64, device=device) for _ in range(depth)]
        self.depth = depth
        self.work = [None] * depth
        if self.cuda:
            self.side = torch.cuda.Stream(device=device)
            self.ready = [torch.cuda.Event() for _ in range(depth)]     # payload written by the evaluation stream
            self.done = [torch.cuda.Event() for _ in range(depth)]      # all-reduce finished on the side stream
            self.used = [False] * depth

    def buffer(self, step, stream=None):
        """Buffer for this step's payload; the evaluation stream first waits for the all-reduce that used it last."""
        k = step % self.depth
        if self.cuda:
            if self.used[k]:
                (stream or self.torch.cuda.current_stream()).wait_event(self.done[k])
        elif self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        return self.buffers[k]

    def submit(self, step, stream=None):
        """Call after the payload of `step` has been enqueued on the evaluation stream."""
        import torch.distributed as dist
        k = step % self.depth
        active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.cuda:
            main = stream or self.torch.cuda.current_stream()
            self.ready[k].record(main)
            self.side.wait_event(self.ready[k])
            if active:
                with self.torch.cuda.stream(self.side):
                    dist.all_reduce(self.buffers[k], op=dist.ReduceOp.SUM, group=self.group)
            self.done[k].record(self.side)
            self.used[k] = True
        elif active:
            self.work[k] = dist.all_reduce(self.buffers[k], op=dist.ReduceOp.SUM, async_op=True, group=self.group)

    def result(self, step, stream=None):
        """The all-reduced payload of `step` (the evaluation stream waits for it; CPU: blocks)."""
        k = step % self.depth
        if self.cuda:
            (stream or self.torch.cuda.current_stream()).wait_event(self.done[k])
        elif self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        return self.buffers[k]
