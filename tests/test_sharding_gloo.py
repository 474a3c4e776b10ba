"""CPU, world_size 2 over gloo: the N>1 host logic — block sharding of loci and the per-step all-reduce of
[sum data lnL | sum genealogy lnL, coal/mig totals] — with the CPU oracle standing in for the kernels.
The all-reduced vector of two half-shards must equal the single-process totals over all loci."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shard = importlib.import_module("g-phocs_b200.shard")
synth = importlib.import_module("g-phocs_b200.synth")


def shard_payload(w, lo, hi):
    """Per-shard payload computed by the oracle (test infrastructure) for loci lo..hi-1 of workload w."""
    from oracle import bindings as ob
    Q, B = len(w.pops["father"]), len(w.pops["band_src"])
    pt, keep = ob.make_poptree(w.pops, w.band_start, w.band_end)
    tc, tm = np.zeros(Q), np.zeros(B)
    tnc, tnm = np.zeros(Q, np.int64), np.zeros(B, np.int64)
    sd = sg = 0.0
    for l in range(lo, hi):
        p0, p1, u0, u1 = int(w.patt_start[l]), int(w.patt_start[l + 1]), int(w.unph_start[l]), int(w.unph_start[l + 1])
        lc = ob.OracleLocus(w.n, w.chars[p0:p1], w.num_phases[p0:p1], w.counts[u0:u1], float(w.rate[l]))
        lc.set_tree(w.father[l], w.left[l], w.right[l], w.age[l], int(w.root[l]))
        sd += lc.compute(0)
        e0, e1 = int(w.ev_start[l]), int(w.ev_start[l + 1])
        _, cs, nc, ms, nm, lnl = ob.oracle_gen_locus(pt, w.pop_start[l], w.ev_type[e0:e1], w.ev_id[e0:e1], w.ev_time[e0:e1])
        sg += lnl
        tc += cs; tnc += nc; tm += ms[:B]; tnm += nm[:B]
    return shard.pack_payload(sd, sg, tc, tnc, tm, tnm)


def _worker(rank, world, port, L, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.generate(synth.config("dip8mig"), L, seed=5)      # same workload on every rank; each takes its block
        lo, hi = shard.shard_range(L, rank, world)
        t = torch.from_numpy(shard_payload(w, lo, hi))
        dist.barrier()
        shard.all_reduce_payload(t)
        if rank == 0:
            np.save(out, t.numpy())
    finally:
        dist.destroy_process_group()


def _pipeline_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipe = shard.PipelinedAllReduce(4, "cpu", depth=2)
        got = []
        for step in range(7):      # step i's payload is reduced while step i+1 is being "computed"
            buf = pipe.buffer(step)
            buf.copy_(torch.tensor([step, rank + 1.0, (rank + 1.0) * step, 1.0], dtype=torch.float64))
            pipe.submit(step)
            if step > 0:
                got.append(pipe.result(step - 1).clone())
        got.append(pipe.result(6).clone())
        if rank == 0:
            np.save(out, torch.stack(got).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_pipelined_allreduce_returns_every_step_once(tmp_path):
    """The side-stream all-reduce of bench.py's step, on gloo: with two buffers in flight every step's payload comes
    back summed over both ranks, in order, one step late."""
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "pipe.npy")
    mp.spawn(_pipeline_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    want = np.array([[2.0 * i, 3.0, 3.0 * i, 2.0] for i in range(7)])
    assert np.array_equal(got, want)


def test_shard_ranges_cover_all_loci_once():
    for L in (1, 7, 100, 100_000):
        for world in (1, 2, 3, 8):
            got = [shard.shard_range(L, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == L
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            assert max(hi - lo for lo, hi in got) == -(-L // world)


def test_payload_roundtrip():
    v = shard.pack_payload(-1.5, 2.5, [1.0, 2.0, 3.0], [4, 5, 6], [0.5], [7])
    u = shard.unpack_payload(v, 3, 1)
    assert u["sum_data_lnl"] == -1.5 and u["sum_gen_lnl"] == 2.5
    assert list(u["total_num_coals"]) == [4, 5, 6] and list(u["total_num_migs"]) == [7]
    assert len(v) == shard.payload_len(3, 1)


@pytest.mark.timeout(300)
def test_two_rank_allreduce_equals_single_process(tmp_path):
    L, world = 60, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, port, L, out), nprocs=world, join=True)
    got = np.load(out)
    w = synth.generate(synth.config("dip8mig"), L, seed=5)
    want = shard_payload(w, 0, L)
    Q, B = len(w.pops["father"]), len(w.pops["band_src"])
    g, e = shard.unpack_payload(got, Q, B), shard.unpack_payload(want, Q, B)
    assert np.array_equal(g["total_num_coals"], e["total_num_coals"]) and g["total_num_coals"].sum() == L * (w.n - 1)
    assert np.array_equal(g["total_num_migs"], e["total_num_migs"])
    assert np.allclose(got, want, rtol=1e-12, atol=0.0)      # fp64 sums differ only by association
