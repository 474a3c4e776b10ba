"""Shared by the GPU tests that compare a device chain with the reference's own chain: runs oracle/_ref/G-PhoCS-ref once
per (shape, loci, iterations, finetunes, priors) and keeps the trace for the other tests of the session."""
import hashlib
import importlib
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "G-PhoCS-ref")
DEVHOST = os.path.join(ROOT, "oracle", "_ref", "G-PhoCS-b200-dev")
synth = importlib.import_module("g-phocs_b200.synth")
CACHE = os.path.join(tempfile.gettempdir(), "gphocs_b200_refchains")

FT = dict(coal_time=0.01, theta=0.3, tau=0.0002, mixing=0.05, mig_time=0.3, mig_rate=0.4)
# Gamma(1, 0.005) on the migration rates: the control-file default Gamma(0.002, 1e-5) puts a spike at the 1e-5 cut-off
# (GPhoCS.c:3159) that chains leave and enter only rarely, which makes means of short chains meaningless
MIG_PRIOR = (1.0, 0.005)


def read_trace(path):
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    names = lines[0].split()
    rows = [[float(x) for x in ln.split()] for ln in lines[1:]]
    width = min(len(r) for r in rows)
    return names, np.array([r[:width] for r in rows])


def batch_se(x, batches=20):
    """Monte-Carlo standard error of the mean of a correlated series by batch means."""
    x = np.asarray(x, float)
    k = len(x) // batches
    means = x[:k * batches].reshape(batches, k).mean(1)
    return means.std(ddof=1) / np.sqrt(batches)


def setup(cfg, L, iters, data_seed=99, finetunes=None, mig_prior=MIG_PRIOR):
    """Directory with the alignment and a control file for `cfg`; returns (dir, model, workload, ctl path)."""
    ft = dict(FT)
    ft.update(finetunes or {})
    key = hashlib.sha1(json.dumps([cfg, L, iters, data_seed, ft, mig_prior], sort_keys=True).encode()).hexdigest()[:16]
    d = os.path.join(CACHE, key)
    os.makedirs(d, exist_ok=True)
    model = synth.config(cfg)
    seq = os.path.join(d, "seqs.txt")
    w = synth.generate(model, L, seed=data_seed, seqfile=seq)
    return d, model, w, ft


def chain(binary, tag, cfg, L, iters, data_seed=99, finetunes=None, mig_prior=MIG_PRIOR, threads=4, seed=4242, timeout=3000):
    """Trace (names, rows) of `binary` on the shape; cached in the session's temp directory."""
    d, model, w, ft = setup(cfg, L, iters, data_seed, finetunes, mig_prior)
    ctl, trace = os.path.join(d, f"{tag}.ctl"), os.path.join(d, f"{tag}.trace")
    done = os.path.join(d, f"{tag}.done")
    if not os.path.exists(done):
        synth.write_control_file(model, ctl, os.path.join(d, "seqs.txt"), trace, iterations=iters, seed=seed,
                                 iterations_per_log=iters, finetunes=ft, mig_prior=mig_prior)
        r = subprocess.run([binary, ctl, "-n", str(threads)], capture_output=True, text=True, timeout=timeout, cwd=d)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        with open(done, "w") as f:
            f.write(r.stdout[-4000:])
    names, rows = read_trace(trace)
    return names, rows, model, w, ft, open(done).read()


def parameter_columns(model, rows):
    """Trace rows -> parameter columns in natural units (print factors of the control files written by synth)."""
    Q, C, B = model.numPops, model.numCurPops, len(model.bands)
    K = 2 * Q - C + B + sum(1 for _ in model.sample_age) + (1 if model.rate_shape > 0 else 0)
    x = rows[:, 1:1 + K].copy()
    x[:, :2 * Q - C] /= 10000.0                                  # tau-theta-print
    x[:, 2 * Q - C:2 * Q - C + B] /= 0.001                       # mig-rate-print
    e0 = 2 * Q - C + B
    x[:, e0:e0 + len(model.sample_age)] /= 10000.0               # estimated sample ages are taus
    return x


def pooled_between_chain_se(group_a, group_b):
    """Standard error of (mean of group a's chain means - mean of group b's) from the spread BETWEEN independent chains:
    each group is [chains][parameters] of per-chain posterior means.  For parameters that mix slowly (the configs[3]
    shape: two reference chains that differ only in their seed disagree by 8 batch-means standard errors on theta_B)
    batch means inside one chain underestimate the error; independent chains do not."""
    a, b = np.asarray(group_a, float), np.asarray(group_b, float)
    dof = max(1, len(a) + len(b) - 2)
    ss = ((a - a.mean(0)) ** 2).sum(0) + ((b - b.mean(0)) ** 2).sum(0)
    return np.sqrt(ss / dof * (1.0 / len(a) + 1.0 / len(b)))


def group_difference(group_a, group_b):
    """(difference of the two groups' means, its standard error) per parameter; each group is a list of chains
    [iterations][parameters].  The error is the LARGER of two estimates: the spread between independent chains
    (pooled_between_chain_se: right when chains mix slowly, but with two or three chains per group it is itself a noisy
    estimate — by chance three chains can agree much better than their own batch-means errors say) and the batch-means
    errors inside the chains combined (right when they mix well)."""
    ma, mb = np.array([c.mean(0) for c in group_a]), np.array([c.mean(0) for c in group_b])
    between = pooled_between_chain_se(ma, mb)

    def within(group):
        se = np.array([[batch_se(c[:, k]) for k in range(c.shape[1])] for c in group])
        return np.sqrt((se ** 2).sum(0)) / len(group)
    return ma.mean(0) - mb.mean(0), np.maximum(between, np.hypot(within(group_a), within(group_b))), ma, mb


REF_SEEDS = (4242, 999, 31337)


def sample_age_columns(model):
    """Columns of parameter_columns() that hold ESTIMATED SAMPLE AGES.  The reference gives them the improper prior
    Gamma(0, 0) ~ 1/x (PopulationTree.c:121), which makes the posterior non-integrable at 0: chains — the reference's as
    much as the device's — wander off towards 0 for long stretches (three reference seeds give 1.5e-5, 2.3e-5 and 3.5e-5
    for tau_B of the configs[4] shape; additive proposals of size 2e-4 are almost never accepted from 1e-8), so their
    "posterior mean" has no Monte-Carlo error worth the name.  The move itself is checked with a proper prior in
    tests/test_gpu_sampler_ancient.py::test_uninformative_data_recovers_the_prior_of_sample_age_and_rates; here these
    columns only have to agree in order of magnitude."""
    Q, C, B = model.numPops, model.numCurPops, len(model.bands)
    e0 = 2 * Q - C + B
    return set(range(e0, e0 + len(model.sample_age)))
