"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgphocs_ref.so).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py

For each small configuration a synthetic alignment + control file is written to a temp dir, the
reference's own pipeline is run on it (readControlFile -> processAlignments -> GetMem ->
performMCMC with a few iterations, GPhoCS.c:146-237), and the per-locus state is dumped through
ref_harness.c: phased patterns as handed to initializeLocusData, genealogies, mutation rates,
data log-likelihoods, node populations, migration nodes, flattened event chains with the
reference's num_lineages, per-locus coal/mig statistics and gtreeLnLikelihood, population
parameters and the running totals.  These are the golden vectors every implementation in this repo
(oracle and CUDA) is checked against on boxes without the reference.
A second fixture kind ("ops_*.npz") records proposal/accept/reject traces driven through the
reference's LocusData API on stand-alone loci (tests/ops.py).
"""
import argparse
import ctypes as C
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# name -> (synth config, loci, mcmc iterations, data seed, missing fraction)
CASES = {
    "sample": ("sample", 24, 25, 1001, 0.0),
    "hap16": ("hap16", 24, 15, 1002, 0.0),
    "dip8mig": ("dip8mig", 24, 25, 1003, 0.05),
    "pop6mig4": ("pop6mig4", 16, 20, 1004, 0.0),
    "ancient": ("ancient", 24, 20, 1005, 0.0),
}


def dump_case(name):
    synth = importlib.import_module("g-phocs_b200.synth")
    from oracle import bindings as ob
    cfg, L, iters, seed, missing = CASES[name]
    model = synth.config(cfg)
    with tempfile.TemporaryDirectory() as tmp:
        seq = os.path.join(tmp, "seqs.txt")
        ctl = os.path.join(tmp, "run.ctl")
        synth.generate(model, L, seed=seed, missing_frac=missing, seqfile=seq)
        synth.write_control_file(model, ctl, seq, os.path.join(tmp, "trace.log"), iterations=iters, seed=4242)
        lib = ob.ref()
        rc = lib.refh_setup(ctl.encode(), 1, 0)
        assert rc == 0, rc
        assert lib.refh_run_mcmc() == 0
        assert lib.refh_check_all() == 1
    L = lib.refh_num_loci()
    n = lib.refh_num_leaves()
    N = 2 * n - 1
    Q, Cn, B = lib.refh_num_pops(), lib.refh_num_cur_pops(), lib.refh_num_bands()
    out = dict(n=n, Q=Q, C=Cn, B=B, L=L, iterations=iters)
    theta, age, sage = np.zeros(Q), np.zeros(Q), np.zeros(Q)
    father, son0, son1, spp = (np.zeros(Q, np.int32) for _ in range(4))
    lib.refh_get_pops(ob.dp(theta), ob.dp(age), ob.dp(sage), ob.ip(father), ob.ip(son0), ob.ip(son1), ob.ip(spp))
    out.update(theta=theta, pop_age=age, sample_age=sage[:Cn], pop_father=father, pop_son0=son0, pop_son1=son1,
               samples_per_pop=spp[:Cn])
    bs, bt = np.zeros(max(B, 1), np.int32), np.zeros(max(B, 1), np.int32)
    br, bst, ben = np.zeros(max(B, 1)), np.zeros(max(B, 1)), np.zeros(max(B, 1))
    lib.refh_get_bands(ob.ip(bs), ob.ip(bt), ob.dp(br), ob.dp(bst), ob.dp(ben))
    out.update(band_src=bs[:B], band_tgt=bt[:B], band_rate=br[:B], band_start=bst[:B], band_end=ben[:B])
    # patterns, as recorded at initializeLocusData
    assert lib.refh_recorded_count() == L
    patt_start, unph_start = [0], [0]
    chars, phases, counts = [], [], []
    for g in range(L):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib.refh_recorded_dims(g, C.byref(a), C.byref(b), C.byref(c))
        assert a.value == n
        ch = np.zeros(b.value * n, np.uint8)
        ph = np.zeros(b.value, np.int32)
        ct = np.zeros(c.value, np.int32)
        lib.refh_recorded_get(g, ch.ctypes.data_as(C.c_char_p), ob.ip(ph), ob.ip(ct))
        chars.append(ch.reshape(b.value, n))
        phases.append(ph)
        counts.append(ct)
        patt_start.append(patt_start[-1] + b.value)
        unph_start.append(unph_start[-1] + c.value)
    out.update(patt_start=np.array(patt_start, np.int64), unph_start=np.array(unph_start, np.int64),
               chars=np.concatenate(chars), num_phases=np.concatenate(phases), counts=np.concatenate(counts))
    # genealogies and likelihoods
    tf, tl, tr, npop = (np.zeros((L, N), np.int32) for _ in range(4))
    ta = np.zeros((L, N))
    root = np.zeros(L, np.int32)
    rate, dlnl, dlnl_full, glnl, glnl_stored = (np.zeros(L) for _ in range(5))
    ev_start, mig_start = [0], [0]
    pop_start = np.zeros((L, Q + 1), np.int32)
    evt, evi, eve, evn = [], [], [], []
    mbr, mba, mtg, msr, mag = [], [], [], [], []
    cs, ms = np.zeros((L, Q)), np.zeros((L, max(B, 1)))
    nc, nm = np.zeros((L, Q), np.int32), np.zeros((L, max(B, 1)), np.int32)
    for g in range(L):
        r_, rt_ = C.c_int(), C.c_double()
        lib.refh_get_tree(g, ob.ip(tf[g]), ob.ip(tl[g]), ob.ip(tr[g]), ob.dp(ta[g]), C.byref(r_), C.byref(rt_))
        root[g], rate[g] = r_.value, rt_.value
        lib.refh_get_node_pops(g, ob.ip(npop[g]))
        dlnl[g] = lib.refh_data_lnl(g)
        glnl_stored[g] = lib.refh_stored_gen_lnl(g)
        E = lib.refh_flatten_events(g, None, None, None, None, None)
        t_, i_, n_ = (np.zeros(E, np.int32) for _ in range(3))
        e_ = np.zeros(E)
        lib.refh_flatten_events(g, ob.ip(pop_start[g]), ob.ip(t_), ob.ip(i_), ob.dp(e_), ob.ip(n_))
        evt.append(t_); evi.append(i_); eve.append(e_); evn.append(n_)
        ev_start.append(ev_start[-1] + E)
        lib.refh_get_stats(g, ob.dp(cs[g]), ob.ip(nc[g]), ob.dp(ms[g]), ob.ip(nm[g]))
        glnl[g] = lib.refh_gen_lnl(g)
        a5 = [np.zeros(10, np.int32) for _ in range(4)]
        ag = np.zeros(10)
        k = lib.refh_get_migs(g, ob.ip(a5[0]), ob.ip(a5[1]), ob.ip(a5[2]), ob.ip(a5[3]), ob.dp(ag))
        mbr.append(a5[0][:k]); mba.append(a5[1][:k]); mtg.append(a5[2][:k]); msr.append(a5[3][:k]); mag.append(ag[:k])
        mig_start.append(mig_start[-1] + k)
    # a full recompute through the reference must reproduce the incrementally maintained values
    for g in range(L):
        dlnl_full[g] = lib.refh_compute_data_lnl(g, 0)
        lib.resetSaved(C.c_void_p(lib.refh_locus(g)))
    tcs, tms = np.zeros(Q), np.zeros(max(B, 1))
    tnc, tnm = np.zeros(Q, np.int32), np.zeros(max(B, 1), np.int32)
    lib.refh_get_total_stats(ob.dp(tcs), ob.ip(tnc), ob.dp(tms), ob.ip(tnm))
    out.update(father=tf, left=tl, right=tr, age=ta, root=root, rate=rate, node_pop=npop, data_lnl=dlnl,
               data_lnl_full=dlnl_full, gen_lnl=glnl, gen_lnl_stored=glnl_stored,
               ev_start=np.array(ev_start, np.int64), pop_start=pop_start, ev_type=np.concatenate(evt),
               ev_id=np.concatenate(evi), ev_time=np.concatenate(eve), ev_lineages=np.concatenate(evn),
               coal_stats=cs, num_coals=nc, mig_stats=ms[:, :B], num_migs=nm[:, :B],
               mig_start=np.array(mig_start, np.int64), mig_branch=np.concatenate(mbr), mig_band=np.concatenate(mba),
               mig_target=np.concatenate(mtg), mig_source=np.concatenate(msr), mig_age=np.concatenate(mag),
               total_coal_stats=tcs, total_num_coals=tnc, total_mig_stats=tms[:B], total_num_migs=tnm[:B],
               total_data_lnl=lib.refh_total_data_lnl())
    np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
    print(f"{name}: L={L} n={n} Q={Q} B={B} sumP={patt_start[-1]} events={ev_start[-1]} migs={mig_start[-1]} "
          f"sum data lnL={dlnl.sum():.6f} sum gen lnL={glnl.sum():.6f}")


def dump_ops():
    """Proposal traces through the reference's stand-alone LocusData API."""
    import ops
    from oracle import bindings as ob
    cases = []
    for seed in range(12):
        rng = np.random.default_rng(7000 + seed)
        n = int(rng.integers(3, 14))
        P = int(rng.integers(1, 20))
        chars, ph, cnt = ops.random_patterns(n, P, rng, diploid_pairs=n // 2 if seed % 2 else 0,
                                             missing=0.1 if seed % 3 == 0 else 0.0)
        leaf_ages = np.where(rng.random(n) < 0.3, 1e-4, 0.0) if seed % 4 == 0 else np.zeros(n)
        f, l, r, a, root = ops.random_tree(n, rng, leaf_ages=leaf_ages)
        rate = 0.5 + rng.random()
        lc = ob.RefLocus(n, chars, ph, cnt, rate)
        lc.set_tree(f, l, r, a, root)
        steps = 40
        tr = ops.run_ops(lc, n, seed, steps, allow_leaf_age=(seed % 4 == 0), rate_moves=(seed % 5 == 0))
        lnls, trees = [], []
        for e in tr:
            if e[0] in ("init", "final-full", "rate"):
                lnls.append([e[1], np.nan])
            else:
                lnls.append([e[1], e[3]])
                trees.append(np.concatenate([e[4], e[5], e[6], [e[8]]]).astype(np.float64).tolist() + e[7].tolist())
        cases.append(dict(seed=seed, n=n, chars=chars, num_phases=ph, counts=cnt, father=f, left=l, right=r, age=a,
                          root=root, rate=rate, steps=steps, leaf=(seed % 4 == 0), ratem=(seed % 5 == 0),
                          lnls=np.array(lnls), trees=np.array(trees)))
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}_{k}"] = np.asarray(v)
    flat["num_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "ops_reference.npz"), **flat)
    print(f"ops: {len(cases)} traces")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--one")
    args = ap.parse_args()
    if args.one == "ops":
        dump_ops()
    elif args.one:
        dump_case(args.one)
    else:
        # the reference keeps its state in globals: one process per case
        for name in list(CASES) + ["ops"]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], check=True,
                           stdout=None if name == "ops" else subprocess.PIPE and None)
