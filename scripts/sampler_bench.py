#!/usr/bin/env python
"""MCMC iterations/s of the device-resident sampler (include/gphocs_b200.h group D) on a synthetic workload.
    python scripts/sampler_bench.py --config hap16 --loci 10000 --iterations 50"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="hap16")
ap.add_argument("--loci", type=int, default=10000)
ap.add_argument("--iterations", type=int, default=50)
ap.add_argument("--stepwise", action="store_true", help="per-node launches instead of the one-launch sweep")
args = ap.parse_args()
w = synth.generate(synth.config(args.config), args.loci, seed=777)
st = gp.LociStore.from_workload(w)
mig = (w.mig_start, w.mig_branch, w.mig_band, w.mig_age) if len(w.pops["band_src"]) else None
sm = gp.Sampler(st, w.pops, w.node_pop, seed=1, migration=mig)
sm.set_stepwise(args.stepwise)
sm.iterate(3, trace=False)
k0 = gp.lib().gphocsKernelLaunchCount()
t0 = time.perf_counter()
tr = sm.iterate(args.iterations)
dt = time.perf_counter() - t0
k1 = gp.lib().gphocsKernelLaunchCount()
v, es, el = sm.check()
s = sm.state()
print(json.dumps({"config": args.config, "loci": args.loci, "leaves": w.n, "iterations": args.iterations, "route": "stepwise" if args.stepwise else "default",
                  "iters_per_s": args.iterations / dt, "ms_per_iter": 1e3 * dt / args.iterations,
                  "kernel_launches_per_iter": (k1 - k0) / args.iterations,
                  "locus_proposals_per_s": (s["proposed"]["coal_time"] + s["proposed"]["spr"]) / (args.iterations + 3) * args.iterations / dt,
                  "accept_rates": {m: float(s["accepted"][m]) / max(1, int(s["proposed"][m])) for m in gp.Sampler.MOVES},
                  "check": {"violations": v, "max_stat_rel_err": es, "max_lnl_rel_err": el},
                  "final_mean_data_lnl_per_locus": float(tr[-1, -2]) / args.loci}))
