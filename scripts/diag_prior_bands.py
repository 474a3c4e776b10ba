"""Prior recovery (every base missing => posterior = prior) on population trees whose migration bands touch ANCESTRAL
populations: a band from a current population into an ancestral one (the E->AB band of configs[3]) and a band between two
ancestral populations (CD->EF there).  Split times are exchangeable Gammas ordered by the tree, so their prior means are
those of order statistics (computed here by plain Monte Carlo).   python scripts/diag_prior_bands.py   (GPU box)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refchain as rc
gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
alpha, beta, ma, mb = 3.0, 3000.0, 3.0, 0.01
rng = np.random.default_rng(1)
g = rng.gamma(alpha, 1.0 / beta, size=(2_000_000, 3))
cases = {
    "current->ancestral": (synth.Model("pb1", [("A", 2), ("B", 2), ("E", 2)], [("AB", "A", "B", 5e-4), ("root", "AB", "E", 1e-3)],
                                       bands=[("E", "AB", 300.0)]),
                           # tau_AB < tau_root: min and max of two draws
                           {"tau_AB": np.minimum(g[:, 0], g[:, 1]).mean(), "tau_root": np.maximum(g[:, 0], g[:, 1]).mean()}),
    "ancestral->ancestral": (synth.Model("pb2", [("A", 2), ("B", 2), ("C", 2), ("D", 2)],
                                         [("AB", "A", "B", 5e-4), ("CD", "C", "D", 6e-4), ("root", "AB", "CD", 1.2e-3)],
                                         bands=[("AB", "CD", 300.0)]),
                             # tau_AB, tau_CD < tau_root: condition on the third being the largest
                             {"tau_AB": g[(g[:, 2] > g[:, 0]) & (g[:, 2] > g[:, 1])][:, 0].mean(),
                              "tau_CD": g[(g[:, 2] > g[:, 0]) & (g[:, 2] > g[:, 1])][:, 1].mean(),
                              "tau_root": g.max(1).mean()}),
}
for name, (m, taus) in cases.items():
    L = 3
    w = synth.generate(m, L, seed=3)
    n = w.n
    chars = np.full((L, n), ord("N"), np.uint8)
    st = gp.LociStore(n, np.arange(L + 1), np.arange(L + 1), chars, np.ones(L, np.int32), np.ones(L, np.int32))
    st.set_trees(w.father, w.left, w.right, w.age, w.root)
    Q, C = m.numPops, m.numCurPops
    sm = gp.Sampler(st, w.pops, w.node_pop, theta_prior=(alpha, beta), tau_prior=(np.full(Q, alpha), np.full(Q, beta)), seed=11,
                    finetunes=(0.01, 0.6, 0.0008, 0.3), migration=(w.mig_start, w.mig_branch, w.mig_band, w.mig_age), mig_prior=(ma, mb),
                    mig_finetunes=(0.3, 0.6))
    sm.iterate(5000, trace=False)
    tr = sm.iterate(int(os.environ.get("ITERS", "150000")))
    assert sm.check()[0] == 0
    print(name, "accepted", {k: int(v) for k, v in sm.state()["accepted"].items() if v})
    names = [f"theta_{p}" for p in m.names] + [f"tau_{p}" for p in m.names[C:]] + ["m"]
    for k, nm in enumerate(names):
        want = alpha / beta if nm.startswith("theta") else (ma / mb if nm == "m" else taus[nm])
        x = tr[:, k]
        se = rc.batch_se(x)
        print("  %-12s mean %.5e  prior %.5e  z %6.2f   (se %.1e)" % (nm, x.mean(), want, (x.mean() - want) / se, se))
    sm.close(); st.close()
