"""configs[3] shape (30 loci): posterior means of every parameter from two reference seeds, the fast-path host program and two seeds of the device sampler side by side, with batch-means standard errors - the run that showed theta_B / theta_AB / m_A->B mixing slowly in the reference itself.
    python scripts/diag_pop6.py      (GPU box; needs oracle/_ref)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refchain as rc
gp = importlib.import_module("g-phocs_b200")
cfg, L, iters = "pop6mig4", 30, 30000
burn = iters // 5
names, ref, model, w, ft, _ = rc.chain(rc.REF, "ref", cfg, L, iters)
names2, ref2, _, _, _, _ = rc.chain(rc.REF, "ref_s2", cfg, L, iters, seed=999)
_, dev, _, _, _, _ = rc.chain(rc.DEVHOST, "dev", cfg, L, iters, threads=2, seed=777)
Q, C, B = model.numPops, model.numCurPops, len(model.bands)
K = 2 * Q - C + B
cols = {"ref": rc.parameter_columns(model, ref)[burn:], "ref2": rc.parameter_columns(model, ref2)[burn:], "devhost": rc.parameter_columns(model, dev)[burn:]}
for seed in (2024, 7):
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=seed, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]),
                    migration=(w.mig_start, w.mig_branch, w.mig_band, w.mig_age), mig_prior=rc.MIG_PRIOR, mig_finetunes=(ft["mig_time"], ft["mig_rate"]))
    tr = sm.iterate(iters)
    cols[f"api{seed}"] = tr[burn:, :K]
    print("state", seed, sm.state()["accepted"], sm.state()["proposed"])
    sm.close(); st.close()
print("%-12s" % "param", *["%22s" % k for k in cols])
for k in range(K):
    print("%-12s" % names[1 + k], *["%12.4e +-%8.1e" % (v[:, k].mean(), rc.batch_se(v[:, k])) for v in cols.values()])
