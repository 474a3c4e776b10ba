"""configs[3] shape (30 loci): posterior means of every parameter from three reference seeds and four seeds of the device sampler side by side, with batch-means standard errors and the z score of the difference of the group means under the pooled between-chain error - the run that showed theta_B / theta_AB / m_A->B mixing slowly in the reference itself.
    python scripts/diag_pop6.py      (GPU box; needs oracle/_ref)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refchain as rc
gp = importlib.import_module("g-phocs_b200")
cfg, L, iters = "pop6mig4", 30, 30000
burn = iters // 5
cols = {}
for seed in rc.REF_SEEDS:
    names, ref, model, w, ft, _ = rc.chain(rc.REF, f"ref{seed}", cfg, L, iters, seed=seed)
    cols[f"ref{seed}"] = rc.parameter_columns(model, ref)[burn:]
Q, C, B = model.numPops, model.numCurPops, len(model.bands)
K = 2 * Q - C + B
for seed in (2024, 7, 99, 4321):
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=seed, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]),
                    migration=(w.mig_start, w.mig_branch, w.mig_band, w.mig_age), mig_prior=rc.MIG_PRIOR, mig_finetunes=(ft["mig_time"], ft["mig_rate"]))
    tr = sm.iterate(iters)
    cols[f"api{seed}"] = tr[burn:, :K]
    print("state", seed, sm.state()["accepted"], sm.state()["proposed"])
    sm.close(); st.close()
print("%-12s" % "param", *["%22s" % k for k in cols], "   z")
refm = np.array([v[:, :K].mean(0) for k, v in cols.items() if k.startswith("ref")])
devm = np.array([v[:, :K].mean(0) for k, v in cols.items() if k.startswith("api")])
se = rc.pooled_between_chain_se(refm, devm)
for k in range(K):
    print("%-12s" % names[1 + k], *["%12.4e +-%8.1e" % (v[:, k].mean(), rc.batch_se(v[:, k])) for v in cols.values()],
          " %6.2f" % ((refm[:, k].mean() - devm[:, k].mean()) / se[k]))
