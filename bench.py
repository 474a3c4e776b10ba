#!/usr/bin/env python
"""bench.py — locus log-likelihood evaluations/s of the G-PhoCS hot path on B200 (BASELINE.json metric).

One *step* = one pass of the hot path over every resident locus: the full Felsenstein data likelihood
(computeLocusDataLikelihood(locus, 0) semantics: every internal conditional vector recomputed and left
resident, phase-averaged root reduction) + the genealogy likelihood (computeGenetreeStats + gtreeLnLikelihood
+ computeTotalStats) + the packed [sum data lnL, sum genealogy lnL, coal/mig totals] vector, all-reduced over
ranks when N > 1.  One evaluation = data + genealogy log-likelihood of one locus.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's own CPU implementation on the host cores

Workload (config.workload): BASELINE.json configs[3] — 100k loci x 1 kb, 6-population tree + 4 migration
bands, 2 unphased diploids per population (24 leaves) — per GPU (weak scaling, loci sharded by rank with no
data-path collective).  Synthetic alignments from g-phocs_b200/synth.py; per-step HBM working set is
several GB (>> 126 MB L2), so no explicit L2 flush is needed.

BASELINE.json's metric has a second half — MCMC iterations/s at 10k / 100k loci — which the JSON line carries inside
the objects the driver keeps: `roofline.mcmc` (device-resident update steps, per configuration: iterations/s, kernel
launches and algorithmic bytes per iteration, fraction of the HBM roofline; the 100k-locus configurations are sharded
over all ranks = strong scaling) and `cpu_baseline.mcmc` (the reference's own OpenMP build on the host cores).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PASSES_PER_STEP = 200      # reference arm: passes over the bounded sample that make up one timed step
METRIC = "locus_lnL_evals_per_sec"
UNIT = "locus-evals/s"


def algorithmic_bytes_data(n, P):
    """bytes_full = 32*P*(2n-1) + 24*(2n-1) + 8*P + 8 per locus (P live phased patterns)."""
    return 32.0 * P * (2 * n - 1) + 24.0 * (2 * n - 1) + 8.0 * P + 8.0


def layout_bytes_data(n, P):
    """What this layout must move per full evaluation of one locus: internal vectors written once (32*P*(n-1));
    leaf masks (8 B per 16 leaves per column), phase/count words, node records + ages, lnL/savedLnL read or written."""
    return 32.0 * P * (n - 1) + 8.0 * P * ((n + 15) // 16) + 8.0 * P + 16.0 * (2 * n - 1) + 24.0


def algorithmic_bytes_gen(E, Q, B):
    """bytes_gen = 32*E + 16*(Q+B) + 8 per locus."""
    return 32.0 * E + 16.0 * (Q + B) + 8.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- reference arm
def reference_sample(cfg, sample_loci, seed, reps, threads, quiet=True):
    """Times the reference's own CPU implementation (oracle/_ref, all host threads) on a bounded sample of
    the workload: OpenMP loop over loci of computeLocusDataLikelihood(locus,0)+resetSaved and
    computeGenetreeStats+gtreeLnLikelihood (ref_harness.c: refh_time_both_once).  Runs in THIS process
    (call it from a subprocess: the reference keeps global state).  Returns per-rep seconds."""
    import ctypes as C
    import tempfile
    synth = importlib.import_module("g-phocs_b200.synth")
    from oracle import bindings as ob
    model = synth.config(cfg)
    with tempfile.TemporaryDirectory() as tmp:
        seq, ctl = os.path.join(tmp, "seqs.txt"), os.path.join(tmp, "run.ctl")
        synth.generate(model, sample_loci, seed=seed, seqfile=seq)
        synth.write_control_file(model, ctl, seq, os.path.join(tmp, "trace.log"), iterations=1, seed=4242)
        lib = ob.ref()
        if quiet:
            sys.stdout.flush()
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)
        try:
            rc = lib.refh_setup(ctl.encode(), threads, 0)
            assert rc == 0, f"reference set-up failed ({rc})"
            lib.refh_init_only()
        finally:
            if quiet:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(devnull)
    lib.refh_set_threads(threads)
    times = []
    sd, sg = C.c_double(), C.c_double()
    # bounded sample: passes per step sized from two calibration passes so that the whole run is ~20 s of CPU work
    t_pass = min(lib.refh_time_both_once(C.byref(sd), C.byref(sg)) for _ in range(2))
    passes = int(max(1, min(PASSES_PER_STEP, round(25.0 / max(reps * t_pass, 1e-9)))))
    for _ in range(reps):
        t = 0.0
        for _p in range(passes):
            t += lib.refh_time_both_once(C.byref(sd), C.byref(sg))
        times.append(t)
    return times, lib.refh_num_loci() * passes, sd.value, sg.value, passes


def port_sample(cfg, sample_loci, seed, reps):
    """Fallback when oracle/_ref is absent: the oracle C port, one thread."""
    synth = importlib.import_module("g-phocs_b200.synth")
    from oracle import bindings as ob
    w = synth.generate(synth.config(cfg), sample_loci, seed=seed)
    loci = []
    for l in range(w.L):
        p0, p1, u0, u1 = int(w.patt_start[l]), int(w.patt_start[l + 1]), int(w.unph_start[l]), int(w.unph_start[l + 1])
        lc = ob.OracleLocus(w.n, w.chars[p0:p1], w.num_phases[p0:p1], w.counts[u0:u1], float(w.rate[l]))
        lc.set_tree(w.father[l], w.left[l], w.right[l], w.age[l], int(w.root[l]))
        loci.append(lc)
    pt, keep = ob.make_poptree(w.pops, w.band_start, w.band_end)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for l, lc in enumerate(loci):
            lc.compute(0)
            lc.reset()
            e0, e1 = int(w.ev_start[l]), int(w.ev_start[l + 1])
            ob.oracle_gen_locus(pt, w.pop_start[l], w.ev_type[e0:e1], w.ev_id[e0:e1], w.ev_time[e0:e1])
        times.append(time.perf_counter() - t0)
    return times, w.L, 0.0, 0.0, 1


MCMC_CONFIG, MCMC_LOCI = "hap16", 10_000     # BASELINE.json configs[1]: 10k loci x 1 kb, 16 haplotypes, 4 populations, no migration


def reference_mcmc(threads, iterations=40, cfg=MCMC_CONFIG):
    """MCMC iterations/s of the reference's own OpenMP build (oracle/_ref/G-PhoCS-ref) on workload `cfg` at MCMC_LOCI loci:
    iterations / (wall - wall of a 1-iteration run), BASELINE.md 3.2.  None when the binary is absent."""
    import tempfile
    binary = os.path.join(ROOT, "oracle", "_ref", "G-PhoCS-ref")
    if not os.path.exists(binary):
        return None
    synth = importlib.import_module("g-phocs_b200.synth")
    model = synth.config(cfg)
    with tempfile.TemporaryDirectory() as tmp:
        seq = os.path.join(tmp, "seqs.txt")
        synth.generate(model, MCMC_LOCI, seed=777, seqfile=seq)
        wall = {}
        for iters in (1, iterations):
            ctl = os.path.join(tmp, f"r{iters}.ctl")
            synth.write_control_file(model, ctl, seq, os.path.join(tmp, f"r{iters}.trace"), iterations=iters, seed=4242,
                                     iterations_per_log=iters)
            t0 = time.perf_counter()
            r = subprocess.run([binary, ctl, "-n", str(threads)], capture_output=True, text=True, cwd=tmp)
            wall[iters] = time.perf_counter() - t0
            if r.returncode != 0:
                return None
    return {"config": f"{cfg}: {MCMC_LOCI} loci", "iterations": iterations, "threads": threads,
            "iters_per_s": (iterations - 1) / max(wall[iterations] - wall[1], 1e-9), "wall_s": wall[iterations], "setup_s": wall[1]}


def device_mcmc(gp, synth, device, total_loci, iterations, rank=0, world=1, cfg=MCMC_CONFIG, peak_gbs=None):
    """MCMC iterations/s of the device-resident update steps (gphocs_b200.h group D) on workload `cfg` with `total_loci`
    loci sharded over `world` ranks (strong scaling); global decisions are taken on NCCL all-reduced sums.
    Iteration-level roofline: the algorithmic bytes of every incremental evaluation of the iteration, 32*P*(2k+1) with k
    conditional vectors recomputed, counted by the kernels themselves, over the iteration time."""
    import torch
    import torch.distributed as dist
    shard = importlib.import_module("g-phocs_b200.shard")
    lo, hi = shard.shard_range(total_loci, rank, world)
    w = synth.generate(synth.config(cfg), hi - lo, seed=777 + rank)
    st = gp.LociStore.from_workload(w, device=device)
    mig = (w.mig_start, w.mig_branch, w.mig_band, w.mig_age) if len(w.pops["band_src"]) else None
    model = synth.config(cfg)
    extra = {}
    if model.sample_age or model.rate_shape > 0:        # configs[4]: estimated sample ages, locus-mut-rate VAR
        st.set_rates([1.0] * w.L)
        extra = dict(estimate_sample_age=[1 if nm in model.sample_age else 0 for nm, _ in model.cur],
                     locus_rate_finetune=0.3 if model.rate_shape > 0 else 0.0)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=1, migration=mig, **extra)
    if world > 1:
        sm.init_nccl(rank, world, locus_offset=lo)      # the library's own communicator: sums stay on the device
    sm.iterate(5, trace=False)
    sm.eval_counters(reset=True)                        # switches the accounting on
    k0 = gp.lib().gphocsKernelLaunchCount()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sm.iterate(iterations, trace=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    evals, ebytes = sm.eval_counters()
    if world > 1:
        t = torch.tensor([dt, -float(evals), -float(ebytes)], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        dt, evals, ebytes = float(t[0].item()), -float(t[1].item()), -float(t[2].item())
    launches = gp.lib().gphocsKernelLaunchCount() - k0
    violations, stat_err, lnl_err = sm.check()
    s = sm.state()
    gbs = ebytes / dt / 1e9
    out = {"config": f"{cfg}: {total_loci} loci over {world} GPU(s)", "iterations": iterations, "iters_per_s": iterations / dt,
           "ms_per_iteration": 1e3 * dt / iterations,
           "kernel_launches_per_iteration": launches / iterations,
           "incremental_evals_per_iteration": evals / iterations,
           "algorithmic_bytes_per_iteration": ebytes / iterations,
           "achieved_gbs_all_gpus": gbs, "frac_of_hbm_peak": (gbs / world / peak_gbs) if peak_gbs else None,
           "accept_rates": {m: round(float(s["accepted"][m]) / max(1, int(s["proposed"][m])), 4) for m in gp.Sampler.MOVES if m != "tau_conflicts"},
           "check": {"violations": int(violations), "max_stat_rel_err_vs_recompute": stat_err,
                     "max_lnl_rel_err_vs_full_recompute": lnl_err}}
    sm.close()
    st.close()
    return out


def evals_at_10k(gp, synth, device, cfg, steps=40):
    """Device-resident evaluations/s of the same step (full data likelihood + genealogy likelihood of every locus) at
    10k loci — BASELINE.json's smaller size.  The working set (tens of MB) would sit in the 126 MB L2, so a 512 MB
    buffer is written between the timed steps."""
    import torch
    L = 10_000
    w = synth.generate(synth.config(cfg), L, seed=4321)
    stream = torch.cuda.Stream(device=device)
    st = gp.LociStore.from_workload(w, device=device, stream=stream.cuda_stream)
    gen = gp.Genealogy(L, w.pops, device=device, stream=stream.cuda_stream)
    gen.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{device}")
    ms = []
    with torch.cuda.stream(stream):
        for i in range(steps + 3):
            flush.fill_(i & 0xff)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            st.evaluate_device(0)
            gen.evaluate_device()
            b.record(stream)
            if i >= 3:
                ms.append((a, b))
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in ms) / len(ms)
    st.close(); gen.close()
    return {"config": workload_name(cfg, L), "evals_per_s": L / (t * 1e-3), "ms_per_step": t, "steps": len(ms),
            "l2": "512 MB written between timed steps (working set would otherwise stay in L2)"}


class quiet_stdout:
    """Sends file descriptor 1 to /dev/null for the duration (C code called through ctypes writes there directly)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)     # what the C side still buffers goes to /dev/null too
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)
        return False


def ingest_bench(gp, synth, device, cfg="dip8mig", loci=10_000, ref_loci=1500):
    """Alignment ingest (header group E): sequence file -> initializeLocusData's arguments.
    Product: text parse on the host threads + k_ingest (two passes) on the device, timed separately.  Reference:
    readSeqFile + processHetPatterns per locus of the compiled reference on the first `ref_loci` loci of the same
    file (its pattern pool makes it slower per locus the more loci it reads, so this flatters it)."""
    import tempfile
    model = synth.config(cfg)
    names = synth.sample_slots(model)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "seqs.txt")
        synth.generate(model, loci, seed=4242, seqfile=path)
        size = os.path.getsize(path)
        gp.Alignment.read(path, names, device=device).close()        # warm-up (page cache, module load)
        t0 = time.perf_counter()
        a = gp.Alignment.read(path, names, device=device)
        wall = time.perf_counter() - t0
        t = a.timings()
        out = {"config": f"{cfg}: {loci} loci x 1 kb, {sum(1 for x in names if x)} samples", "file_bytes": size,
               "loci_per_s_e2e": loci / wall, "file_GBps_e2e": size / wall / 1e9, "seconds": {k: t[k] for k in ("parse_s", "h2d_s", "kernel_s", "d2h_s")},
               "kernel_loci_per_s": loci / t["kernel_s"],
               "kernel_GBps": 2 * t["raw_bytes"] / t["kernel_s"] / 1e9,   # both passes read every symbol row once
               "patterns": int(a.U), "phased_patterns": int(a.P)}
        a.close()
        try:
            from oracle import bindings as ob
            from oracle import ingest as oi
            if ob.have_ref():
                t0 = time.perf_counter()
                with quiet_stdout():      # the reference reports its progress on stdout; this script prints one JSON line
                    r = oi.reference_ingest(path, names, ref_loci)
                dt = time.perf_counter() - t0
                out["reference"] = {"loci": len(r), "loci_per_s": len(r) / dt, "cores": 1,
                                    "sample": f"first {ref_loci} loci of the same file, single thread (the reference ingest is serial)"}
        except Exception as e:   # reported, never required
            out["reference"] = {"unavailable": str(e)[:200]}
    return out


def cpu_baseline_subprocess(cfg, sample_loci, reps):
    """cpu_baseline leg: run the reference sample in a child process, parse its JSON."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(reps), "--warmup", "1",
           "--config", cfg, "--sample-loci", str(sample_loci), "--gpus", "1", "--with-mcmc"]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = [x for x in out.stdout.splitlines() if x.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:   # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {e}"[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bindings as ob
    threads = os.cpu_count() or 1
    t_all0 = time.perf_counter()
    if ob.have_ref():
        kind = "reference"
        times, L, sd, sg, passes = reference_sample(args.config, args.sample_loci, 4242, args.steps + args.warmup, threads)
    else:
        kind, threads = "port", 1
        times, L, sd, sg, passes = port_sample(args.config, min(args.sample_loci, 500), 4242, args.steps + args.warmup)
    timed = times[args.warmup:]
    total = float(sum(timed))
    value = L * len(timed) / total
    sample = (f"{L // passes} loci of workload {args.config} (same generator, seed 4242), "
              f"{len(timed)} steps x {passes} passes of "
              f"computeLocusDataLikelihood(locus,0)+resetSaved and computeGenetreeStats+gtreeLnLikelihood over all loci, "
              f"OpenMP static schedule, {threads} threads; wall {time.perf_counter() - t_all0:.1f}s incl. ingest")
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
    if args.with_mcmc:
        cb["mcmc"] = {"configs1_hap16_10k": reference_mcmc(os.cpu_count() or 1),
                      "configs2_dip8mig_10k": reference_mcmc(os.cpu_count() or 1, cfg="dip8mig")}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(args.config, args.loci, args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def common_config(cfg, L, gpus):
    """`config` of the JSON line — the same object for both arms (what differs, e.g. the reference arm's bounded
    sample, is described in `cpu_baseline.sample`)."""
    return {"workload": workload_name(cfg, L) + " per GPU", "loci_per_gpu": L, "n_gpus": gpus, "host_cores": os.cpu_count() or 1,
            "l2": "per-step HBM working set of several GB >> 126 MB L2 (no flush needed)",
            "parallelism": f"loci sharded over {gpus} GPU(s), one all-reduce of a (2+2Q+2B)-double vector per step on a side stream"}


def workload_name(cfg, L):
    synth = importlib.import_module("g-phocs_b200.synth")
    m = synth.config(cfg)
    return (f"{cfg}: {L} loci x {m.sites} bp, {m.numLeaves} leaves ({'unphased diploids' if m.diploid else 'haploid'}), "
            f"{m.numCurPops}-population tree, {len(m.bands)} migration bands")


# --------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    gp = importlib.import_module("g-phocs_b200")
    synth = importlib.import_module("g-phocs_b200.synth")
    shard = importlib.import_module("g-phocs_b200.shard")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgphocs_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = None
        if args.nccl_high_priority:   # NCCL's kernels on a high-priority stream: they start when issued, not after the
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)   # evaluation kernel's last wave of CTAs
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)

    L = args.loci
    model = synth.config(args.config)
    w = synth.generate(model, L, seed=1000 + 17 * rank)     # each rank owns its own shard of loci
    n, Q, B = w.n, model.numPops, len(model.bands)
    P = np.diff(w.patt_start).astype(np.float64)
    E = np.diff(w.ev_start).astype(np.float64)
    bytes_data = float(algorithmic_bytes_data(n, P).sum())
    bytes_gen = float(algorithmic_bytes_gen(E, Q, B).sum())
    bytes_layout = float(layout_bytes_data(n, P).sum())

    stream = torch.cuda.Stream(device=dev)
    st = gp.LociStore.from_workload(w, device=local_rank, stream=stream.cuda_stream)
    gen = gp.Genealogy(L, w.pops, device=local_rank, stream=stream.cuda_stream)
    gen.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    lib = gp.lib()
    host_threads = lib.gphocsSetHostThreads(max(1, (os.cpu_count() or 1) // world))   # torchrun pins OMP_NUM_THREADS=1
    V = 1 + 2 * Q + 2 * B
    assert 1 + V == shard.payload_len(Q, B)
    # [sum data lnL | sum gen lnL, totals...] of step i is summed over ranks on a side stream while step i+1 computes
    pipe = shard.PipelinedAllReduce(1 + V, dev, depth=args.pipe_depth)
    step_no = [0]

    sp_main = C.c_void_p(stream.cuda_stream)

    def step(record=None, pipe=pipe):
        """device-resident pass; `record` = (e0, e1, e2) timing events to record around the two evaluations.  Everything
        here names its stream explicitly (no stream context per step: at eight ranks the host side of a step is what
        the 0.5 ms of kernels have to hide)."""
        i = step_no[0]
        step_no[0] += 1
        if record is not None:
            record[0].record(stream)
        _, dsum = st.evaluate_device(0)
        if record is not None:
            record[1].record(stream)
        _, dtot, v = gen.evaluate_device()
        if record is not None:
            record[2].record(stream)
        payload = pipe.buffer(i, stream)
        base = payload.data_ptr()
        lib.gphocsCopyDeviceAsync(C.c_void_p(base), C.c_void_p(dsum), 8, sp_main)
        lib.gphocsCopyDeviceAsync(C.c_void_p(base + 8), C.c_void_p(dtot), 8 * V, sp_main)
        pipe.submit(i, stream)                # the only cross-GPU traffic: < 1 KB per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.gphocsKernelLaunchCount()
    rec = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]   # created up front
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        t_start.record(stream)
    for k_step in range(args.steps):
        step(rec[k_step])
    with torch.cuda.stream(stream):
        payload = pipe.result(step_no[0] - 1, stream)     # the last step's sums are in before the clock stops
        t_end.record(stream)
    barrier()
    launches = lib.gphocsKernelLaunchCount() - launches0
    ms_total = t_start.elapsed_time(t_end)
    ms_data = sum(a.elapsed_time(b) for a, b, _ in rec) / len(rec)
    ms_gen = sum(b.elapsed_time(c) for _, b, c in rec) / len(rec)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    total_data_lnl, total_gen_lnl = float(payload[0].item()), float(payload[1].item())
    if args.step_only:   # development: the device-resident step alone (scaling experiments), a few variants in one process
        def timed(pipe_v, nsteps):
            for _ in range(5):
                step(None, pipe_v)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            with torch.cuda.stream(stream):
                a.record(stream)
            for _ in range(nsteps):
                step(None, pipe_v)
            with torch.cuda.stream(stream):
                pipe_v.result(step_no[0] - 1, stream)
                b.record(stream)
            barrier()
            ms = a.elapsed_time(b)
            if world > 1:
                tt = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
            return ms / nsteps
        out = {"step_only": True, "n_gpus": world, "steps": args.steps, "ms_per_step": ms_total / args.steps,
               "ms_data": ms_data, "ms_gen": ms_gen, "variants": {}}
        groups = {"default": None}
        if world > 1:
            groups["nccl_high_priority"] = dist.new_group(pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
        for gname, grp in groups.items():
            for depth in (2, 4):
                pv = shard.PipelinedAllReduce(1 + V, dev, depth=depth, group=grp)
                for nsteps in (20, 200):
                    out["variants"][f"{gname}/depth{depth}/steps{nsteps}"] = timed(pv, nsteps)
        if rank == 0:
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: the same pass through the C ABI with HOST buffers (H2D of genealogies + event snapshots, D2H of
    # per-locus log-likelihoods and totals inside the timed region)
    lnl_host = gp.pinned_like(np.zeros(L))
    hw = {k: gp.pinned_like(getattr(w, k)) for k in ("father", "left", "right", "age", "root", "ev_start", "pop_start",
                                                       "ev_type", "ev_id", "ev_time")}   # inputs in page-locked host memory
    # the same inputs in the library's wire formats (16-bit topology triples, 16-bit event codes and chain offsets,
    # 32-bit event offsets): what a host that flattens its genealogies for the GPU writes instead of int32 arrays
    es32, ps16, code16 = gp.pack_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id)
    hp = {"topo": gp.pinned_like(gp.pack_trees(w.father, w.left, w.right)), "es": gp.pinned_like(es32),
          "ps": gp.pinned_like(ps16), "code": gp.pinned_like(code16)}
    e2e_steps = max(3, min(args.steps, 20))
    # bytes that cross PCIe per step, counted from the buffers the library copies.  Packed route: int16 topology
    # triples, fp64 ages, roots; uint16 event codes, fp64 elapsed times, uint16 chain offsets, int32 event offsets.
    # int32 route: the page-locked arrays as they are (int32 topology / event types / ids / chain offsets, int64
    # event offsets).  Back: per-locus data lnL + its sum, genealogy lnL + totals.
    nN = w.father.shape[1]
    h2d = L * nN * (3 * 2 + 8) + 4 * L + int(E.sum()) * 10 + L * (Q + 1) * 2 + 4 * (L + 1)
    h2d_int32 = L * nN * (3 * 4 + 8) + 4 * L + int(E.sum()) * 16 + L * (Q + 1) * 4 + 8 * (L + 1)
    d2h = 8 * L + 8 + 8 * L + 8 * V

    # The host calls are the asynchronous ones of the C ABI (gphocsStoreEvaluateDevice, gphocsGenEvaluateDevice,
    # gphocsCopyDeviceAsync): the data-likelihood kernel and its read-back run while the event snapshot crosses
    # PCIe; everything is back in host memory before the step ends.
    lib.gphocsGenSetStream(gen.h, None)        # the genealogy object on a stream of its own for this section
    host_sums = gp.pinned_like(np.zeros(2 + V))

    def e2e_step(packed=True):
        if packed:
            st.set_trees_packed(hp["topo"], hw["age"], hw["root"])
        else:
            st.set_trees(hw["father"], hw["left"], hw["right"], hw["age"], hw["root"])
        dlnl, dsum = st.evaluate_device(0)
        sp = C.c_void_p(stream.cuda_stream)
        lib.gphocsCopyDeviceAsync(C.c_void_p(lnl_host.ctypes.data), C.c_void_p(dlnl), 8 * L, sp)
        lib.gphocsCopyDeviceAsync(C.c_void_p(host_sums.ctypes.data), C.c_void_p(dsum), 8, sp)
        if packed:
            gen.set_events_packed(hp["es"], hp["ps"], hp["code"], hw["ev_time"])
        else:
            gen.set_events(hw["ev_start"], hw["pop_start"], hw["ev_type"], hw["ev_id"], hw["ev_time"])
        r = gen.evaluate(per_locus_stats=False)      # per-locus genealogy lnL + totals, back on the host
        stream.synchronize()
        return float(host_sums[0]), r["sum_lnl"]

    def e2e_run(packed):
        e2e_step(packed)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sd_, sg_ = e2e_step(packed)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * L * e2e_steps / dt, sd_, sg_

    e2e_int32_value, sdata, sgen = e2e_run(False)
    # the end-to-end pass computes the same numbers as the resident one (this rank's sums before the all-reduce)
    if world == 1:
        assert abs(sdata - total_data_lnl) <= 1e-9 * abs(total_data_lnl), (sdata, total_data_lnl)
        assert abs(sgen - total_gen_lnl) <= 1e-9 * abs(total_gen_lnl), (sgen, total_gen_lnl)
    e2e_full_value, sdata, sgen = e2e_run(True)
    if world == 1:
        assert abs(sdata - total_data_lnl) <= 1e-9 * abs(total_data_lnl), (sdata, total_data_lnl)
        assert abs(sgen - total_gen_lnl) <= 1e-9 * abs(total_gen_lnl), (sgen, total_gen_lnl)

    # ---- e2e, delta route (the headline): an MCMC-shaped host step.  Between two steps of a chain a locus' genealogy
    # changes in a few nodes and one or two event chains, so the host ships what changed — 24-byte edit records
    # (resetSaved + adjustGenNodeAge per locus, gphocsStoreApplyOpsAsync) and the new elapsed times of one population's
    # chain per locus (gphocsGenRecalcAsync) — and the device evaluates everything from scratch as in the resident
    # step: full data likelihood of every locus, full genealogy likelihood of every locus, per-locus values and totals
    # back in page-locked host memory before the step ends.  tests/test_gpu_parity.py::test_delta_upload_step_equals_
    # the_full_upload: the state these deltas leave on the device is bit for bit the one a full upload leaves.
    node = n + 2
    nodeArr = np.full(L, node)
    recs = []
    for f in (1.001, 1.0):          # two edit sets, alternating, so the genealogies stay where they are on average
        r_ = np.zeros(2 * L, gp.OP_DTYPE)
        r_["locus"] = np.repeat(np.arange(L), 2)
        r_["type"] = np.tile([gp.OP_COMMIT, gp.OP_ADJUST_AGE], L)
        r_["a"] = np.tile([0, node], L)
        fa = w.father[np.arange(L), nodeArr]
        bound = np.where(fa >= 0, w.age[np.arange(L), np.maximum(fa, 0)], np.inf)     # stay below the father
        r_["x"][1::2] = np.minimum(w.age[np.arange(L), nodeArr] * f, bound)
        recs.append(gp.pinned_like(r_))
    chain_pop = (np.arange(L) % Q).astype(np.int32)
    a0 = (w.ev_start[:-1] + w.pop_start[np.arange(L), chain_pop]).astype(np.int64)
    b0 = (w.ev_start[:-1] + w.pop_start[np.arange(L), chain_pop + 1]).astype(np.int64)
    starts = np.zeros(L + 1, np.int32); starts[1:] = np.cumsum(b0 - a0)
    idx = np.concatenate([np.arange(x, y) for x, y in zip(a0, b0)])
    pin_loc, pin_pop, pin_starts = gp.pinned_like(np.arange(L, dtype=np.int32)), gp.pinned_like(chain_pop), gp.pinned_like(starts)
    pin_times = [gp.pinned_like(w.ev_time[idx] * f) for f in (0.999, 1.0)]
    h2d_delta = 2 * L * 24 + 4 * L + 4 * L + 4 * (L + 1) + 8 * len(idx)
    st.set_trees_packed(hp["topo"], hw["age"], hw["root"])          # both objects back at the workload's state
    gen.set_events_packed(hp["es"], hp["ps"], hp["code"], hw["ev_time"])
    st.evaluate_device(0); gen.evaluate(per_locus_stats=False)
    stream.synchronize()

    def delta_step(i):
        st.apply_ops_async(recs[i & 1])
        dlnl, dsum = st.evaluate_device(0)
        sp = C.c_void_p(stream.cuda_stream)
        lib.gphocsCopyDeviceAsync(C.c_void_p(lnl_host.ctypes.data), C.c_void_p(dlnl), 8 * L, sp)
        lib.gphocsCopyDeviceAsync(C.c_void_p(host_sums.ctypes.data), C.c_void_p(dsum), 8, sp)
        gen.recalc_async(pin_loc, pin_pop, pin_starts, pin_times[i & 1])
        r = gen.evaluate(per_locus_stats=False)
        stream.synchronize()
        return float(host_sums[0]), r["sum_lnl"]

    for i in range(2):
        delta_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        sd_, sg_ = delta_step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st.sync(); gen.sync()                                           # reports records / chains the device refused
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * L * e2e_steps / dt
    assert np.isfinite(sd_) and np.isfinite(sg_)
    if world == 1:      # (with several ranks total_data_lnl is the all-reduced sum, sd_ this rank's)
        assert abs(sd_ - total_data_lnl) < 1e-3 * abs(total_data_lnl), (sd_, total_data_lnl)
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    st.set_trees_packed(hp["topo"], hw["age"], hw["root"])
    st.evaluate(0, out=lnl_host)

    # ---- MCMC-style cycle (extra): one node-age proposal per locus -> incremental evaluation -> accept/reject
    node = n + 2
    prop = gp.make_ops(np.arange(L), gp.OP_ADJUST_AGE, a=node, x=w.age[:, node] * 1.001)
    rej = gp.make_ops(np.arange(L), gp.OP_REVERT)
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    for _ in range(2):
        st.apply_ops(prop); st.evaluate(1, out=lnl_host); st.apply_ops(rej)
    cyc = 5
    t0 = time.perf_counter()
    for _ in range(cyc):
        st.apply_ops(prop); st.evaluate(1, out=lnl_host); st.apply_ops(rej)
    cyc_s = time.perf_counter() - t0
    with torch.cuda.stream(stream):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        st.apply_ops(prop)
        a.record(stream)
        st.evaluate_device(1)
        b.record(stream)
    torch.cuda.synchronize()
    inc_ms = a.elapsed_time(b)
    st.apply_ops(rej)
    # MCMC iterations/s of the device-resident steps (BASELINE.json's second metric).  The 100k / 50k-locus
    # configurations are sharded over all ranks (strong scaling: total loci fixed); the 10k-locus ones run at N = 1.
    peak, peak_src = measured_peak()
    mcmc = {}
    if not args.no_mcmc:
        mcmc["configs3_pop6mig4_100k_sharded"] = device_mcmc(gp, synth, local_rank, 100_000, 10, rank, world, cfg="pop6mig4", peak_gbs=peak)
        mcmc["hap16_100k_sharded"] = device_mcmc(gp, synth, local_rank, 100_000, 40, rank, world, peak_gbs=peak)
        mcmc["configs4_ancient_50k_sharded"] = device_mcmc(gp, synth, local_rank, 50_000, 20, rank, world, cfg="ancient", peak_gbs=peak)
        if world == 1:
            mcmc["configs1_hap16_10k"] = device_mcmc(gp, synth, local_rank, MCMC_LOCI, 100, peak_gbs=peak)
            mcmc["configs2_dip8mig_10k"] = device_mcmc(gp, synth, local_rank, 10_000, 50, cfg="dip8mig", peak_gbs=peak)
    small = None
    if world == 1 and not args.no_mcmc:
        small = {"configs1_hap16_10k": evals_at_10k(gp, synth, local_rank, "hap16"),
                 "configs2_dip8mig_10k": evals_at_10k(gp, synth, local_rank, "dip8mig")}
    clocks = sampler.stop() if rank == 0 else None
    ingest = ingest_bench(gp, synth, local_rank) if (rank == 0 and not args.no_cpu_baseline) else None

    if rank == 0:
        survey_gbs = bytes_data / (ms_data * 1e-3) / 1e9
        layout_gbs = bytes_layout / (ms_data * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("workload") == f"{args.config}:{L}":
                traffic = tj.get("k_eval_dram_bytes_per_launch")
                traffic_src = ("stored figure, NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one k_eval launch "
                               "from the ncu --set full capture named in profiles/traffic.json")
        value = world * L * args.steps / (ms_total * 1e-3)
        cb = cpu_baseline_subprocess(args.config, args.sample_loci, 3) if (world == 1 and not args.no_cpu_baseline) else None
        mcmc_head = mcmc.get("configs3_pop6mig4_100k_sharded")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": common_config(args.config, L, world),
            # the fraction reported is the one on the bytes this layout really moves (== the ncu DRAM traffic); the
            # SURVEY 8d formula, which also counts leaf vectors that are stored here as 4-bit masks, is kept beside it
            "roofline": {"bound": "hbm", "kernel": "k_eval (full data-likelihood evaluation, all loci)",
                         "achieved": layout_gbs, "peak": peak, "unit": "GB/s", "frac": layout_gbs / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_launch": bytes_layout,
                         "bytes_formula": "this layout's minimum per locus: 32*P*(n-1) written + 8*P*ceil(n/16) + 8*P + 16*(2n-1) + 24 read",
                         "survey_formula": {"bytes_per_launch": bytes_data, "achieved": survey_gbs, "frac": survey_gbs / peak,
                                            "formula": "SURVEY.md 8d: 32*P*(2n-1) + 24*(2n-1) + 8*P + 8 per locus (counts 32-byte leaf "
                                                       "vectors that do not exist here: a layout win, not bandwidth)"},
                         "ms_per_launch": ms_data, "mean_phased_patterns": float(P.mean()), "mean_events": float(E.mean()),
                         "genealogy_kernel_ms": ms_gen, "genealogy_achieved_gbs": bytes_gen / (ms_gen * 1e-3) / 1e9,
                         "mcmc": mcmc, "evals_per_s_at_10k_loci": small},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_delta), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "host_threads_per_rank": host_threads,
                    "route": "delta upload of an MCMC-shaped step: gphocsStoreApplyOpsAsync (resetSaved + adjustGenNodeAge records, "
                             "one node per locus) + gphocsGenRecalcAsync (new elapsed times of one chain per locus), then the full data "
                             "and genealogy evaluation of every locus; per-locus values and totals back in host memory",
                    "full_upload": {"value": e2e_full_value, "h2d_bytes_per_step": int(h2d),
                                    "route": "every genealogy and the whole event snapshot re-sent each step: gphocsStoreSetTreesPacked + "
                                             "gphocsGenSetEventsPacked (16-bit topology and event codes on the wire)"},
                    "full_upload_int32": {"value": e2e_int32_value, "h2d_bytes_per_step": int(h2d_int32),
                                          "route": "gphocsStoreSetTrees + gphocsGenSetEvents (int32 arrays copied as they are)"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "mcmc_iters_per_s": mcmc_head["iters_per_s"] if mcmc_head else None,
            "mcmc_config": mcmc_head["config"] if mcmc_head else None,
            "extra": {"sum_data_lnl": total_data_lnl, "sum_gen_lnl": total_gen_lnl,
                      "mcmc_cycle_proposals_per_sec_e2e": world * L * cyc / cyc_s,
                      "incremental_eval_ms_device": inc_ms,
                      "incremental_evals_per_sec_device": L / (inc_ms * 1e-3),
                      "device_bytes_store": st.device_bytes, "ingest": ingest},
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    st.close()
    gen.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="pop6mig4")
    ap.add_argument("--loci", type=int, default=100_000, help="loci per GPU")
    ap.add_argument("--sample-loci", type=int, default=20000,
                    help="loci in the reference arm's bounded sample (working set >> host last-level cache)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--step-only", action="store_true", help="development: time the device-resident step, print it, stop")
    ap.add_argument("--pipe-depth", type=int, default=2, help="steps the per-step all-reduce may lag behind the evaluation")
    ap.add_argument("--nccl-high-priority", type=int, default=0, help="1: NCCL's stream gets high priority")
    ap.add_argument("--no-mcmc", action="store_true", help="skip the device-resident MCMC extras (development runs)")
    ap.add_argument("--with-mcmc", action="store_true", help="reference arm: also time the reference's MCMC iterations/s")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
