"""GPU: device-resident MCMC update steps with migration bands (sampler_mig.cuh).  Same three kinds of evidence as
tests/test_gpu_sampler.py: statistics against the oracle, prior recovery with uninformative data (now including the
migration rate), posterior means against the reference's own chain, and the checkAll invariants after iterating."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
from oracle import bindings as ob  # noqa: E402
from test_gpu_dropin import REF, read_trace  # noqa: E402
from test_gpu_sampler import batch_se  # noqa: E402


def migration_of(w):
    return (w.mig_start, w.mig_branch, w.mig_band, w.mig_age)


@pytest.mark.parametrize("cfg", ["dip8mig", "pop6mig4", "sample"])
def test_segment_statistics_match_the_oracle(cfg):
    """coal / mig statistics computed from branch segments equal the event-chain statistics of the oracle
    (recalcStats, patch.c:2387-2513) on genealogies with migration events."""
    w = synth.generate(synth.config(cfg), 300, seed=12)
    assert len(w.mig_age) > 0
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, migration=migration_of(w), seed=3)
    got = sm.stats()
    pt, keep = ob.make_poptree(w.pops, w.band_start, w.band_end)
    for l in range(w.L):
        e0, e1 = int(w.ev_start[l]), int(w.ev_start[l + 1])
        _, cs, nc, ms, nm, lnl = ob.oracle_gen_locus(pt, w.pop_start[l], w.ev_type[e0:e1], w.ev_id[e0:e1], w.ev_time[e0:e1])
        assert np.array_equal(got["num_coals"][l], nc)
        assert np.array_equal(got["num_migs"][l], nm[:sm.B])
        assert np.allclose(got["coal"][l], cs, rtol=1e-10, atol=1e-15)
        assert np.allclose(got["mig"][l], ms[:sm.B], rtol=1e-10, atol=1e-15)
    assert sm.check()[0] == 0
    sm.close(); st.close()


def test_state_stays_consistent_with_migration():
    w = synth.generate(synth.config("dip8mig"), 400, seed=31)
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, migration=migration_of(w), seed=5)
    tr = sm.iterate(25)
    assert np.all(np.isfinite(tr))
    v, es, el = sm.check()
    assert v == 0, v
    assert es < 1e-9 and el < 1e-9, (es, el)
    s = sm.state()
    for move in ("coal_time", "spr", "theta", "mig_rate", "mig_time"):
        assert 0 < s["accepted"][move] <= s["proposed"][move], (move, s)
    sm.close(); st.close()


def test_uninformative_data_recovers_the_prior_with_migration():
    """Two current populations + root, one band A -> B; every base missing => posterior = prior: thetas, the root split
    time and the migration rate come out Gamma(alpha, beta)."""
    m = synth.Model("prior_mig", [("A", 3), ("B", 3)], [("root", "A", "B", 1e-3)], bands=[("A", "B", 300.0)])
    L = 3
    w = synth.generate(m, L, seed=3)
    n = w.n
    chars = np.full((L, n), ord("N"), np.uint8)
    st = gp.LociStore(n, np.arange(L + 1), np.arange(L + 1), chars, np.ones(L, np.int32), np.ones(L, np.int32))
    st.set_trees(w.father, w.left, w.right, w.age, w.root)
    alpha, beta = 3.0, 3000.0
    ma, mb = 3.0, 0.01            # migration rate ~ Gamma(3, 0.01): mean 300
    Q = 3
    sm = gp.Sampler(st, w.pops, w.node_pop, theta_prior=(alpha, beta), tau_prior=(np.full(Q, alpha), np.full(Q, beta)), seed=11,
                    finetunes=(0.01, 0.6, 0.0008, 0.3), migration=migration_of(w), mig_prior=(ma, mb), mig_finetunes=(0.3, 0.6))
    sm.iterate(3000, trace=False)
    tr = sm.iterate(60000)
    assert sm.check()[0] == 0
    for col, name, mean, sd in [(0, "theta_A", alpha / beta, np.sqrt(alpha) / beta), (1, "theta_B", alpha / beta, np.sqrt(alpha) / beta),
                                (2, "theta_root", alpha / beta, np.sqrt(alpha) / beta), (3, "tau_root", alpha / beta, np.sqrt(alpha) / beta),
                                (4, "m_A->B", ma / mb, np.sqrt(ma) / mb)]:
        x = tr[:, col]
        se = batch_se(x)
        assert abs(x.mean() - mean) < 3.5 * se + 0.01 * mean, (name, x.mean(), mean, se)
        assert abs(x.std() - sd) < 0.15 * sd, (name, x.std(), sd)
    s = sm.state()
    assert s["accepted"]["mig_time"] > 0 and s["accepted"]["mig_rate"] > 0
    sm.close(); st.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/G-PhoCS-ref not built")
@pytest.mark.parametrize("cfg,L", [("sample", 60), ("dip8mig", 40)])
def test_posterior_means_match_the_reference_chain_with_migration(cfg, L):
    """Unphased diploids, migration bands — the sample shape and configs[2]: posterior means of every theta, tau and
    migration rate from the device chain against the reference's own chain on the same alignment (same priors and
    finetunes), within 3 Monte-Carlo standard errors (+ 1 %: the batch-means error estimate is itself uncertain)."""
    import refchain as rc
    iters = 30000
    burn = iters // 5
    names, ref, model, w, ft, _ = rc.chain(rc.REF, "ref", cfg, L, iters)
    Q, C, B = model.numPops, model.numCurPops, len(model.bands)
    K = 2 * Q - C + B
    ref = rc.parameter_columns(model, ref)[burn:]
    st = gp.LociStore.from_workload(w)
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=2024, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]),
                    migration=migration_of(w), mig_prior=rc.MIG_PRIOR, mig_finetunes=(ft["mig_time"], ft["mig_rate"]))
    tr = sm.iterate(iters)[burn:, :K]
    assert sm.check()[0] == 0
    for k in range(K):
        a, b = ref[:, k], tr[:, k]
        se = np.hypot(batch_se(a), batch_se(b))
        assert abs(a.mean() - b.mean()) < 3.0 * se + 0.01 * abs(a.mean()), (names[1 + k], a.mean(), b.mean(), se)
    sm.close(); st.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/G-PhoCS-ref not built")
def test_posterior_means_match_the_reference_on_the_headline_shape():
    """configs[3] shape (6 populations, 4 bands, 24 leaves — the shape of the bench headline) at 30 loci.  Some of its
    parameters (theta_B, theta_AB, m_A->B: the populations joined by a band) mix slowly in the reference itself: two
    reference chains that differ only in their seed disagree by 8 batch-means standard errors.  The Monte-Carlo error
    is therefore taken from INDEPENDENT chains — three seeds of the reference, three of the device sampler — and the
    means of the two groups must agree within 3 standard errors of their difference (refchain.group_difference: the
    larger of the between-chain and the batch-means estimate) + 2 %."""
    import refchain as rc
    cfg, L, iters = "pop6mig4", 30, 16000
    burn = iters // 4
    ref_chains, model, w, ft, names = [], None, None, None, None
    for seed in rc.REF_SEEDS:
        names, ref, model, w, ft, _ = rc.chain(rc.REF, f"ref_{seed}", cfg, L, iters, seed=seed)
        ref_chains.append(rc.parameter_columns(model, ref)[burn:])
    Q, C, B = model.numPops, model.numCurPops, len(model.bands)
    K = 2 * Q - C + B
    dev_chains = []
    for seed in (2024, 7, 90210):
        st = gp.LociStore.from_workload(w)
        sm = gp.Sampler(st, w.pops, w.node_pop, seed=seed, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]),
                        migration=migration_of(w), mig_prior=rc.MIG_PRIOR, mig_finetunes=(ft["mig_time"], ft["mig_rate"]))
        tr = sm.iterate(iters)[burn:, :K]
        assert sm.check()[0] == 0
        dev_chains.append(tr)
        sm.close(); st.close()
    diff, se, ref_means, dev_means = rc.group_difference([c[:, :K] for c in ref_chains], dev_chains)
    for k in range(K):
        a = ref_means[:, k].mean()
        assert abs(diff[k]) < 3.0 * se[k] + 0.02 * abs(a), (names[1 + k], a, diff[k], se[k], ref_means[:, k], dev_means[:, k])
