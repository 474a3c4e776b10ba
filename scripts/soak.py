#!/usr/bin/env python
"""Long chains of the device-resident steps with the checkAll invariants tested along the way (rare-event hunt):
    python scripts/soak.py [iterations] [loci]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ok = True
for cfg in ("sample", "hap16", "dip8mig", "pop6mig4", "ancient"):
    model = synth.config(cfg)
    w = synth.generate(model, L, seed=2024)
    st = gp.LociStore.from_workload(w)
    mig = (w.mig_start, w.mig_branch, w.mig_band, w.mig_age) if len(w.pops["band_src"]) else None
    extra = {}
    if model.sample_age or model.rate_shape > 0:
        st.set_rates(np.ones(w.L))
        extra = dict(estimate_sample_age=[1 if nm in model.sample_age else 0 for nm, _ in model.cur], locus_rate_finetune=0.3)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=99, migration=mig, mig_prior=(1.0, 0.005), finetunes=(0.01, 0.04, 0.00002, 0.003), **extra)
    t0 = time.time()
    worst = (0, 0.0, 0.0)
    for k in range(0, iters, 500):
        tr = sm.iterate(min(500, iters - k))
        v, es, el = sm.check()
        worst = (max(worst[0], v), max(worst[1], es), max(worst[2], el))
        if v or es > 1e-8 or el > 1e-8 or not np.all(np.isfinite(tr)):
            ok = False
            print(json.dumps({"config": cfg, "FAILED_after": k + 500, "violations": v, "stat_err": es, "lnl_err": el}))
            break
    s = sm.state()
    print(json.dumps({"config": cfg, "loci": L, "iterations": iters, "seconds": round(time.time() - t0, 1), "violations": worst[0],
                      "max_stat_err": worst[1], "max_lnl_err": worst[2],
                      "accept": {m: round(float(s["accepted"][m]) / max(1, int(s["proposed"][m])), 3) for m in gp.Sampler.MOVES},
                      "tau_conflicts": int(s["proposed"]["tau_conflicts"]), "last_row": [float(x) for x in tr[-1, -2:]]}))
    sm.close(); st.close()
sys.exit(0 if ok else 1)
