"""GPU: device-resident MCMC update steps (include/gphocs_b200.h group D).

The chains use their own random streams, so parity with the reference is statistical:
  * with uninformative data (every base missing) the chain must sample the PRIOR: thetas and the root split time
    come out Gamma(alpha, beta) — a check of every move's acceptance ratio that needs no reference at all;
  * on real synthetic alignments the posterior means of every theta and tau must agree with the reference's own
    chain (oracle/_ref/G-PhoCS-ref, same control file) within Monte-Carlo error;
  * after any number of iterations the device state passes the reference's checkAll invariants (patch.c:2745):
    incrementally maintained statistics and log-likelihoods equal a recomputation from scratch."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
from test_gpu_dropin import REF, read_trace  # noqa: E402


def batch_se(x, batches=20):
    """Monte-Carlo standard error of the mean of a correlated series by batch means."""
    x = np.asarray(x, float)
    k = len(x) // batches
    means = x[:k * batches].reshape(batches, k).mean(1)
    return means.std(ddof=1) / np.sqrt(batches)


def test_state_stays_consistent_on_real_data():
    w = synth.generate(synth.config("hap16"), 400, seed=31)
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=5)
    v, es, el = sm.check()
    assert v == 0 and es < 1e-10 and el < 1e-12
    tr = sm.iterate(30)
    assert np.all(np.isfinite(tr))
    v, es, el = sm.check()
    assert v == 0, v
    assert es < 1e-9 and el < 1e-9, (es, el)
    s = sm.state()
    for move in ("coal_time", "spr", "theta", "tau"):
        assert 0 < s["accepted"][move] <= s["proposed"][move], (move, s)
    # the genealogy log-density in the trace equals the kernels' from-scratch value for the downloaded state
    node_pop = sm.download()
    f, l, r, a, root = st.get_trees()
    assert node_pop.shape == f.shape and np.all(a[np.arange(len(root)), root] >= s["tau"][len(w.pops["father"]) - 1] - 1e-12)
    sm.close()
    st.close()


def test_two_nodes_per_lane_path_stays_consistent():
    """24 leaves (47 nodes): the migration-free kernels hold two nodes per lane; six populations, diploid data."""
    m = synth.config("pop6mig4")
    m.bands = []
    w = synth.generate(m, 300, seed=8)
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=9)
    tr = sm.iterate(12)
    assert np.all(np.isfinite(tr))
    v, es, el = sm.check()
    assert v == 0 and es < 1e-9 and el < 1e-9, (v, es, el)
    s = sm.state()
    assert 0 < s["accepted"]["spr"] < s["proposed"]["spr"] and 0 < s["accepted"]["coal_time"]
    sm.close()
    st.close()


@pytest.mark.parametrize("leaves", [40, 72])
def test_many_leaves_per_locus_stay_consistent(leaves):
    """79 and 143 nodes per genealogy: four and thirteen nodes per lane in the migration-free kernels, several leaf
    words per column in k_eval."""
    k = leaves // 2
    m = synth.Model("wide", [("A", k), ("B", leaves - k)], [("root", "A", "B", 1e-3)])
    w = synth.generate(m, 60, seed=8)
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=9)
    tr = sm.iterate(6)
    assert np.all(np.isfinite(tr))
    v, es, el = sm.check()
    assert v == 0 and es < 1e-9 and el < 1e-9, (v, es, el)
    s = sm.state()
    assert 0 < s["accepted"]["spr"] < s["proposed"]["spr"] and 0 < s["accepted"]["coal_time"]
    sm.close()
    st.close()


@pytest.mark.parametrize("cfg,L", [("hap16", 300), ("ancient", 200), ("pop6nomig", 150), ("mid", 80), ("dense", 24), ("wide32", 60), ("tiny", 50)])
def test_sweep_routes_give_the_same_chain(cfg, L):
    """The one-launch sweep (sweep_kernels.cuh: a CTA keeps its batch of loci for both per-locus sweeps) and the stepwise
    route (a proposal launch and a k_eval launch per node) use the same random streams and the same arithmetic:
    identical traces, statistics, genealogies, population assignments, log-likelihoods and conditional vectors —
    with a tenth of the launches.  Shapes: configs[1]; configs[4] (sample ages, locus rates); 24 leaves / 6
    populations (two nodes per lane on the stepwise route); denser patterns (a few loci per CTA batch, loci wider
    than a warp); dense patterns (every locus wider than a CTA: walked in chunks, root terms through HBM scratch);
    32 leaves (the largest the sweep takes); 3 leaves (fewer nodes than team threads)."""
    if cfg == "pop6nomig":
        base = synth.config("pop6mig4")
        model = synth.Model("pop6nomig", base.cur, base.anc, diploid=base.diploid)
    elif cfg == "mid":          # theta 5e-3: tens of patterns per locus, a few loci per CTA batch
        model = synth.config("hap16")
        model.theta = 5e-3
        model.anc = [(a, b, c, t * 5) for a, b, c, t in model.anc]
    elif cfg == "wide32":
        model = synth.Model("wide32", [("A", 16), ("B", 16)], [("root", "A", "B", 1e-3)])
    elif cfg == "tiny":
        model = synth.Model("tiny", [("A", 2), ("B", 1)], [("root", "A", "B", 1e-3)])
    else:
        model = synth.config(cfg)
    w = synth.generate(model, L, seed=41)
    assert not len(w.pops["band_src"])
    out = {}
    for route in ("stepwise", "sweep"):
        st = gp.LociStore.from_workload(w)
        extra = {}
        if model.sample_age:
            st.set_rates(np.ones(w.L))
            extra = dict(estimate_sample_age=[1 if nm in model.sample_age else 0 for nm, _ in model.cur], locus_rate_finetune=0.3)
        sm = gp.Sampler(st, w.pops, w.node_pop, seed=77, **extra)
        sm.set_stepwise(route == "stepwise")
        sm.eval_counters(reset=True)           # switches the roofline accounting on
        k0 = gp.lib().gphocsKernelLaunchCount()
        tr = sm.iterate(12)
        launches = gp.lib().gphocsKernelLaunchCount() - k0
        counters = sm.eval_counters()
        assert sm.check()[0] == 0
        stats = sm.stats()
        state = sm.state()
        state["eval_counters"] = counters
        node_pop = sm.download()
        trees = st.get_trees()
        clvs = [st.clv(l, st.n + k, P=int(w.patt_start[l + 1] - w.patt_start[l])) for l in (0, L // 2, L - 1) for k in (0, st.n - 2)]
        out[route] = (tr, stats, node_pop, trees, st.lnl(), launches, clvs, state)
        sm.close(); st.close()
    a, b = out["stepwise"], out["sweep"]
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1]["coal"], b[1]["coal"]) and np.array_equal(a[1]["num_coals"], b[1]["num_coals"])
    assert np.array_equal(a[2], b[2])
    for x, y in zip(a[3], b[3]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[4], b[4])
    for x, y in zip(a[6], b[6]):
        assert np.array_equal(x, y)
    for move in ("coal_time", "spr"):
        assert a[7]["accepted"][move] == b[7]["accepted"][move] and a[7]["proposed"][move] == b[7]["proposed"][move], move
    # both routes count the same incremental evaluations and the same algorithmic bytes, 32 * P * (2k + 1) each
    evals, nbytes = a[7]["eval_counters"]
    assert a[7]["eval_counters"] == b[7]["eval_counters"] and evals > 0
    P = np.diff(w.patt_start)
    assert 32 * P.min() * 3 * evals <= nbytes <= 32 * P.max() * (2 * (w.n - 1) + 1) * evals
    assert b[5] < a[5] / 2, (b[5], a[5])


def test_uninformative_data_recovers_the_prior():
    """Two current populations (3 haploids each) + root; every base missing => likelihood 1 => posterior = prior.
    theta_A, theta_B, theta_root ~ Gamma(3, 3000) and tau_root ~ Gamma(3, 3000) marginally (sum over genealogies of
    P(G | theta, tau) is 1), whatever the number of loci."""
    m = synth.Model("prior", [("A", 3), ("B", 3)], [("root", "A", "B", 1e-3)])
    L = 3
    w = synth.generate(m, L, seed=3)
    n = w.n
    chars = np.full((L, n), ord("N"), np.uint8)
    st = gp.LociStore(n, np.arange(L + 1), np.arange(L + 1), chars, np.ones(L, np.int32), np.ones(L, np.int32))
    st.set_trees(w.father, w.left, w.right, w.age, w.root)
    alpha, beta = 3.0, 3000.0
    Q = 3
    sm = gp.Sampler(st, w.pops, w.node_pop, theta_prior=(alpha, beta), tau_prior=(np.full(Q, alpha), np.full(Q, beta)), seed=11,
                    finetunes=(0.01, 0.6, 0.0008, 0.3))
    sm.iterate(2000, trace=False)                       # burn-in
    tr = sm.iterate(40000)
    assert sm.check()[0] == 0
    mean, sd = alpha / beta, np.sqrt(alpha) / beta
    for col, name in [(0, "theta_A"), (1, "theta_B"), (2, "theta_root"), (3, "tau_root")]:
        x = tr[:, col]
        se = batch_se(x)
        assert abs(x.mean() - mean) < 3.5 * se + 0.01 * mean, (name, x.mean(), mean, se)
        assert abs(x.std() - sd) < 0.12 * sd, (name, x.std(), sd)
    assert np.all(tr[:, -2] == 0.0)                     # data log-likelihood is identically zero
    sm.close()
    st.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/G-PhoCS-ref not built")
def test_posterior_means_match_the_reference_chain():
    """BASELINE.json configs[1] shape (16 haplotypes, 4 populations, no migration) at 60 loci: posterior means of
    every theta and tau from the device chain against the reference's own chain on the same alignment, within
    3 Monte-Carlo standard errors (+ 1 %: the batch-means error estimate is itself uncertain)."""
    import refchain as rc
    L, iters = 60, 30000
    burn = iters // 5
    names, ref, model, w, ft, _ = rc.chain(rc.REF, "ref", "hap16", L, iters)
    Q, C = model.numPops, model.numCurPops
    ref = rc.parameter_columns(model, ref)[burn:]
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=2024, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]))
    tr = sm.iterate(iters)[burn:, :2 * Q - C]
    assert sm.check()[0] == 0
    for k in range(2 * Q - C):
        a, b = ref[:, k], tr[:, k]
        se = np.hypot(batch_se(a), batch_se(b))
        assert abs(a.mean() - b.mean()) < 3.0 * se + 0.01 * abs(a.mean()), (names[1 + k], a.mean(), b.mean(), se)
    sm.close()
    st.close()
