"""configs[4] shape (50 loci): the same table for three reference seeds, two runs of the fast-path host program and two seeds of the device sampler - the run that showed the estimated sample age (improper prior) wandering in every chain.
    python scripts/diag_ancient.py      (GPU box; needs oracle/_ref)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refchain as rc
gp = importlib.import_module("g-phocs_b200")
cfg, L, iters = "ancient", 50, 30000
burn = iters // 5
names, ref, model, w, ft, _ = rc.chain(rc.REF, "ref", cfg, L, iters)
names2, ref2, _, _, _, _ = rc.chain(rc.REF, "ref_s2", cfg, L, iters, seed=999)
_, dev, _, _, _, _ = rc.chain(rc.DEVHOST, "dev", cfg, L, iters, threads=2, seed=777)
_, dev2, _, _, _, _ = rc.chain(rc.DEVHOST, "dev_s2", cfg, L, iters, threads=2, seed=4711)
_, ref3, _, _, _, _ = rc.chain(rc.REF, "ref_s3", cfg, L, iters, seed=31337)
Q, C, B = model.numPops, model.numCurPops, len(model.bands)
K = 2 * Q - C + B + 2
cols = {"ref": rc.parameter_columns(model, ref)[burn:], "ref2": rc.parameter_columns(model, ref2)[burn:], "ref3": rc.parameter_columns(model, ref3)[burn:], "devhost": rc.parameter_columns(model, dev)[burn:], "devhost2": rc.parameter_columns(model, dev2)[burn:]}
for seed in (2024, 7):
    st = gp.LociStore.from_workload(w)
    st.set_rates(np.ones(w.L))
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=seed, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]),
                    estimate_sample_age=[1 if nm in model.sample_age else 0 for nm, _ in model.cur], locus_rate_finetune=0.3)
    tr = sm.iterate(iters)
    cols[f"api{seed}"] = tr[burn:, :K]
    print("state", seed, sm.state()["accepted"], sm.state()["proposed"])
    sm.close(); st.close()
print("%-12s" % "param", *["%22s" % k for k in cols])
for k in range(K):
    print("%-12s" % names[1 + k], *["%12.4e +-%8.1e" % (v[:, k].mean(), rc.batch_se(v[:, k])) for v in cols.values()])
