"""GPU, 2 ranks (skipped on a single-GPU box): loci sharded over ranks, every global decision of the device-resident
sampler taken on all-reduced sums (NCCL), so both ranks must hold identical theta / tau after every iteration and
each rank's shard must pass the checkAll invariants.  Run by `gpurun --gpus 2 -- python -m pytest tests -m gpu`."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, mode, cfg="hap16"):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    gp = importlib.import_module("g-phocs_b200")
    synth = importlib.import_module("g-phocs_b200.synth")
    shard = importlib.import_module("g-phocs_b200.shard")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        L = 600
        lo, hi = shard.shard_range(L, rank, world)
        model = synth.config(cfg)
        w = synth.generate(model, hi - lo, seed=100 + rank)      # this rank's block of loci
        st = gp.LociStore.from_workload(w, device=rank)
        extra = {}
        if model.rate_shape > 0:                                  # configs[4]: estimated sample ages, locus-mut-rate VAR
            st.set_rates(np.ones(w.L))
            extra = dict(estimate_sample_age=[1 if nm in model.sample_age else 0 for nm, _ in model.cur], locus_rate_finetune=0.3)
        sm = gp.Sampler(st, w.pops, w.node_pop, seed=7, **extra)

        def all_reduce(v):
            buf = torch.from_numpy(v.copy()).to(f"cuda:{rank}")
            dist.all_reduce(buf)
            v[:] = buf.cpu().numpy()
        if mode == "hook":
            sm.set_all_reduce(all_reduce, locus_offset=lo)
        else:                                             # the library's own NCCL communicator (gphocsSamplerInitNccl)
            sm.init_nccl(rank, world, locus_offset=lo)
        tr = sm.iterate(25)
        v, es, el = sm.check()
        np.save(os.path.join(out_dir, f"trace{rank}.npy"), tr)
        np.save(os.path.join(out_dir, f"check{rank}.npy"), np.array([v, es, el]))
        sm.download()                                       # brings the store's host mirror (rates included) up to date
        np.save(os.path.join(out_dir, f"rates{rank}.npy"), st.get_rates())
        sm.close(); st.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["hook", "nccl"])
def test_two_ranks_keep_identical_parameters(tmp_path, mode):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), mode), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "trace0.npy"), np.load(tmp_path / "trace1.npy")
    keep = os.path.join(os.path.dirname(str(tmp_path)), "multi_trace_hook.npy")
    if mode == "hook":
        np.save(keep, t0)
    elif os.path.exists(keep):                            # summing two ranks is order-free: both routes give the same bits
        assert np.array_equal(np.load(keep), t0)
    assert np.array_equal(t0, t1)                      # thetas, taus and the all-reduced log-likelihood sums
    assert np.all(np.isfinite(t0)) and len(np.unique(t0[:, 0])) > 1
    for r in range(2):
        v, es, el = np.load(tmp_path / f"check{r}.npy")
        assert v == 0 and el < 1e-9


@pytest.mark.parametrize("mode", ["hook", "nccl"])
def test_locus_rates_move_between_ranks(tmp_path, mode):
    """UpdateLocusRate (GPhoCS.c:4598-4675) moves rate between any two loci of the data set: pairs are formed over the
    global locus index, so rate crosses GPUs.  The sum over ALL ranks stays at the number of loci while each rank's own
    sum drifts away from its locus count; both ranks still agree on every global parameter and on Variance-Mut."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), mode, "ancient"), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "trace0.npy"), np.load(tmp_path / "trace1.npy")
    assert np.array_equal(t0, t1) and np.all(np.isfinite(t0))
    r0, r1 = np.load(tmp_path / "rates0.npy"), np.load(tmp_path / "rates1.npy")
    L = len(r0) + len(r1)
    assert abs(r0.sum() + r1.sum() - L) < 1e-9 * L          # one simplex over all ranks ...
    assert abs(r0.sum() - len(r0)) > 1e-6                   # ... not one per rank
    assert r0.min() > 0 and r1.min() > 0 and r0.std() > 0
    var_mut = np.sqrt(np.mean((np.concatenate([r0, r1]) - 1.0) ** 2))
    assert abs(t0[-1, -3] - var_mut) < 1e-12                # the trace's Variance-Mut is over all ranks' loci
    for r in range(2):
        v, es, el = np.load(tmp_path / f"check{r}.npy")
        assert v == 0 and el < 1e-9
