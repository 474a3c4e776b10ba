#!/usr/bin/env python
"""MCMC iterations/s of the reference host: its own CPU build (oracle/_ref/G-PhoCS-ref, OpenMP) against the same
host objects linked to libgphocs_b200.so (oracle/_ref/G-PhoCS-b200: parallel regions run as fibers, likelihoods
batched on the GPU).  Same alignment, control file and seed; iterations/s = iterations / (wall - wall of a
1-iteration run), as BASELINE.md §3.2 prescribes.  Also checks that the two traces coincide.

    python scripts/mcmc_bench.py --config hap16 --loci 10000 --iterations 30 [--threads N]
"""
import argparse, importlib, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
synth = importlib.import_module("g-phocs_b200.synth")
from test_gpu_dropin import read_trace, REF, B200


def run(binary, ctl, threads, env=None):
    t0 = time.perf_counter()
    r = subprocess.run([binary, ctl, "-n", str(threads)], capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(f"{binary} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    for ln in r.stderr.splitlines():
        if ln.startswith("gphocs_b200:"):
            print(ln, file=sys.stderr)
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="hap16")
    ap.add_argument("--loci", type=int, default=10000)
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--skip-ref", action="store_true")
    args = ap.parse_args()
    model = synth.config(args.config)
    tmp = tempfile.mkdtemp(prefix="gphocs_mcmc_")
    seq = os.path.join(tmp, "seqs.txt")
    synth.generate(model, args.loci, seed=777, seqfile=seq)
    out = {"config": args.config, "loci": args.loci, "iterations": args.iterations, "threads": args.threads}
    traces = {}
    for tag, binary in (("reference_cpu", REF), ("b200_fibers", B200)):
        if tag == "reference_cpu" and args.skip_ref:
            continue
        wall = {}
        for iters in (1, args.iterations):
            ctl, trace = os.path.join(tmp, f"{tag}_{iters}.ctl"), os.path.join(tmp, f"{tag}_{iters}.trace")
            synth.write_control_file(model, ctl, seq, trace, iterations=iters, seed=4242, iterations_per_log=max(iters, 1))
            wall[iters] = run(binary, ctl, args.threads)
            if iters > 1:
                traces[tag] = read_trace(trace)[1]
        out[tag] = {"wall_s": wall[args.iterations], "setup_s": wall[1],
                    "iters_per_s": (args.iterations - 1) / max(wall[args.iterations] - wall[1], 1e-9)}
    if len(traces) == 2:
        a, b = traces["reference_cpu"], traces["b200_fibers"]
        same = np.all(np.isclose(a, b, rtol=1e-6, atol=1e-9), axis=1)
        out["traces_identical_rows"] = int(same.sum())
        out["speedup"] = out["b200_fibers"]["iters_per_s"] / out["reference_cpu"]["iters_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
