"""GPU: the fast-path host (oracle/_ref/G-PhoCS-b200-dev = the unmodified reference host through initializeMCMC +
g-phocs_b200/host/gphocs_device_mcmc.c) runs the SAME control file as the reference program and
writes the same trace format; its posterior means agree with the reference's own chain within Monte-Carlo error on
all five shapes of BASELINE.json `configs`, and it is much faster than the reference on the box's host cores."""
import os
import time

import numpy as np
import pytest
import torch

import refchain as rc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not (os.path.exists(rc.REF) and os.path.exists(rc.DEVHOST)), reason="oracle/_ref binaries not built")]

SHAPES = [("hap16", 60, 30000), ("sample", 60, 30000), ("dip8mig", 40, 30000), ("ancient", 50, 30000)]


@pytest.mark.parametrize("cfg,L,iters", SHAPES)
def test_control_file_run_matches_reference_posterior(cfg, L, iters):
    burn = iters // 5
    names_r, ref, model, w, ft, _ = rc.chain(rc.REF, "ref", cfg, L, iters)
    names_d, dev, _, _, _, log = rc.chain(rc.DEVHOST, "dev", cfg, L, iters, threads=2, seed=777)
    assert names_r == names_d                      # same trace header, literally
    assert ref.shape == dev.shape and dev.shape[0] == iters
    assert "MCMC done" in log and "inconsistency" not in log
    a, b = rc.parameter_columns(model, ref)[burn:], rc.parameter_columns(model, dev)[burn:]
    loose = rc.sample_age_columns(model)          # improper prior: see refchain.sample_age_columns
    for k in range(a.shape[1]):
        if k in loose:
            assert 0.25 < a[:, k].mean() / b[:, k].mean() < 4.0, (names_r[1 + k], a[:, k].mean(), b[:, k].mean())
            continue
        se = np.hypot(rc.batch_se(a[:, k]), rc.batch_se(b[:, k]))
        assert abs(a[:, k].mean() - b[:, k].mean()) < 3.0 * se + 0.01 * abs(a[:, k].mean()), \
            (names_r[1 + k], a[:, k].mean(), b[:, k].mean(), se)
    # the two log-likelihood columns (mean full lnL per locus, data lnL) describe the same posterior
    for col in (-2, -1):
        x, y = ref[burn:, col], dev[burn:, col]
        se = np.hypot(rc.batch_se(x), rc.batch_se(y))
        assert abs(x.mean() - y.mean()) < 3.0 * se + 2e-3 * abs(x.mean()), (names_r[col], x.mean(), y.mean(), se)


def test_control_file_run_matches_reference_posterior_on_the_headline_shape():
    """configs[3] shape at 30 loci.  Parameters of the populations a band joins mix slowly in the reference itself (see
    tests/test_gpu_sampler_mig.py), so the Monte-Carlo error comes from independent chains: three seeds of the reference
    program, two of the fast-path host on the same control file."""
    cfg, L, iters = "pop6mig4", 30, 16000
    burn = iters // 4
    ref_chains, dev_chains = [], []
    for seed in rc.REF_SEEDS:
        names_r, ref, model, _, _, _ = rc.chain(rc.REF, f"ref_{seed}", cfg, L, iters, seed=seed)
        ref_chains.append(rc.parameter_columns(model, ref)[burn:])
    for seed in (777, 4711):
        names_d, dev, _, _, _, log = rc.chain(rc.DEVHOST, f"dev_{seed}", cfg, L, iters, threads=2, seed=seed)
        assert names_r == names_d and dev.shape[0] == iters and "MCMC done" in log
        dev_chains.append(rc.parameter_columns(model, dev)[burn:])
    # 3 standard errors of the difference of the group means (the larger of the between-chain and the batch-means
    # estimate: with five chains the former alone is too noisy an estimate) + 2 %
    diff, se, ref_means, dev_means = rc.group_difference(ref_chains, dev_chains)
    for k in range(ref_means.shape[1]):
        a = ref_means[:, k].mean()
        assert abs(diff[k]) < 3.0 * se[k] + 0.02 * abs(a), (names_r[1 + k], a, diff[k], se[k], ref_means[:, k], dev_means[:, k])


def test_control_file_run_is_faster_than_the_reference(tmp_path):
    """configs[1] at 2000 loci, 60 iterations: wall time of the whole program, ingest and set-up included."""
    t = {}
    for tag, binary, threads in (("ref", rc.REF, os.cpu_count() or 1), ("dev", rc.DEVHOST, os.cpu_count() or 1)):
        t0 = time.perf_counter()
        rc.chain(binary, tag + "_speed", "hap16", 2000, 60, threads=threads)
        t[tag] = time.perf_counter() - t0
    assert t["dev"] < t["ref"], t
