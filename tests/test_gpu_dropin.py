"""GPU drop-in test: the UNMODIFIED reference host (GPhoCS.c, patch.c, ... compiled from /root/reference in
the build container, oracle/_ref/G-PhoCS-b200) linked against the product library instead of
LocusDataLikelihood.o.  Every data-likelihood evaluation of its MCMC (computeLocusDataLikelihood,
scaleAllNodeAges, checkLocusDataLikelihood inside checkAll, ...) runs on the GPU through the reference's own
LocusData call surface.  Its seeded chain must reproduce the trace of the reference's own CPU build
(oracle/_ref/G-PhoCS-ref) — the host, its RNG and its proposal code are the same program, so the traces agree
to printed precision unless a last-ulp difference in a log-likelihood flips an accept/reject;
the test therefore compares the first iterations line by line and the rest through posterior means."""
import importlib
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "G-PhoCS-ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "G-PhoCS-b200")
synth = importlib.import_module("g-phocs_b200.synth")


def read_trace(path):
    """Trace file -> (header names, float matrix).  Own parser: readTrace mis-parses v1.3.2 traces."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    names = lines[0].split()
    rows = [[float(x) for x in ln.split()] for ln in lines[1:]]
    width = min(len(r) for r in rows)
    return names, np.array([r[:width] for r in rows])


def run_chain(binary, tmp, tag, model, loci, iterations, threads):
    seq = os.path.join(tmp, "seqs.txt")
    if not os.path.exists(seq):
        synth.generate(model, loci, seed=321, seqfile=seq)
    ctl, trace = os.path.join(tmp, f"{tag}.ctl"), os.path.join(tmp, f"{tag}.trace")
    synth.write_control_file(model, ctl, seq, trace, iterations=iterations, seed=4242, iterations_per_log=10)
    r = subprocess.run([binary, ctl, "-n", str(threads)], capture_output=True, text=True, timeout=1500, cwd=tmp)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Fatal Error" not in r.stdout and "Inconsistent" not in r.stdout, r.stdout[-3000:]
    return read_trace(trace)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("cfg,loci,iterations", [("sample", 100, 60), ("ancient", 60, 40), ("dip8mig", 40, 30)])
def test_reference_host_on_gpu_library_reproduces_reference_chain(tmp_path, cfg, loci, iterations):
    model = synth.config(cfg)
    tmp = str(tmp_path)
    names_r, ref = run_chain(REF, tmp, "ref", model, loci, iterations, threads=1)
    names_g, gpu = run_chain(B200, tmp, "b200", model, loci, iterations, threads=2)
    assert names_r == names_g and ref.shape == gpu.shape and ref.shape[0] == iterations
    # identical host + RNG: the chains coincide until (if ever) an accept/reject flips on a last-ulp difference
    same = np.all(np.isclose(ref, gpu, rtol=1e-6, atol=1e-9), axis=1)
    first_diff = int(np.argmin(same)) if not same.all() else iterations
    assert first_diff >= min(10, iterations), (first_diff, ref[first_diff], gpu[first_diff])
    # posterior means of every parameter column (theta, tau, m, ...) over the chain
    assert np.allclose(ref.mean(0), gpu.mean(0), rtol=0.05, atol=1e-6)
    # the log-likelihood columns agree to north_star's tolerance wherever the chains coincide
    k = first_diff
    assert np.allclose(ref[:k, -2:], gpu[:k, -2:], rtol=1e-6)
