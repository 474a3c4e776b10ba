"""GPU: device-resident update steps for BASELINE.json configs[4] — ancient samples (UpdateSampleAge, GPhoCS.c:4006)
and per-locus rate variation (UpdateLocusRate, GPhoCS.c:4598).  Evidence as in tests/test_gpu_sampler.py: invariants
after iterating, prior recovery with uninformative data, posterior means against the reference's own chain."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
from test_gpu_dropin import REF, read_trace  # noqa: E402
from test_gpu_sampler import batch_se  # noqa: E402


def estimated(model):
    return np.array([1 if nm in model.sample_age else 0 for nm, _ in model.cur], np.int32)


def test_state_stays_consistent_with_ancient_samples_and_rates():
    model = synth.config("ancient")
    w = synth.generate(model, 401, seed=17)          # odd number of loci: one locus sits out every rate round
    st = gp.LociStore.from_workload(w)
    st.set_rates(np.ones(w.L))
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=5, estimate_sample_age=estimated(model), locus_rate_finetune=0.3,
                    finetunes=(0.01, 0.04, 0.00002, 0.003))
    assert sm.width == 2 * sm.Q - sm.C + 2 + 2
    tr = sm.iterate(40)
    assert np.all(np.isfinite(tr))
    v, es, el = sm.check()
    assert v == 0, v
    assert es < 1e-9 and el < 1e-9, (es, el)
    s = sm.state()
    for move in ("coal_time", "spr", "theta", "tau", "sample_age", "locus_rate"):
        assert 0 < s["accepted"][move] <= s["proposed"][move], (move, s)
    b = [nm for nm, _ in model.cur].index("B")
    assert s["tau"][b] != model.sample_age["B"] and np.all(s["tau"][:sm.C][estimated(model) == 0] == 0.0)
    sm.download()                                    # device state -> host mirror
    leaves_b = np.asarray(w.node_pop).reshape(w.L, -1)[:, :w.n] == b
    ages = st.get_trees()[3][:, :w.n]
    assert np.all(ages[leaves_b] == s["tau"][b]) and np.all(ages[~leaves_b] == 0.0)
    rates = st.get_rates()
    assert abs(rates.sum() - w.L) < 1e-9 * w.L and rates.min() > 0 and rates.std() > 0
    assert abs(tr[-1, -3] - np.sqrt(np.mean((rates - 1) ** 2))) < 1e-12
    sm.close(); st.close()


def test_uninformative_data_recovers_the_prior_of_sample_age_and_rates():
    """Every base missing => posterior = prior.  The sample age s of A and the root split time tau have independent
    Gamma priors restricted to s < tau (moments by rejection sampling); the locus rates / L are Dirichlet(alpha):
    Var(rate) = (L - 1) / (alpha L + 1)."""
    m = synth.Model("prior_anc", [("A", 3), ("B", 3)], [("root", "A", "B", 1e-3)], sample_age={"A": 1e-4})
    L = 3
    w = synth.generate(m, L, seed=3)
    n = w.n
    chars = np.full((L, n), ord("N"), np.uint8)
    st = gp.LociStore(n, np.arange(L + 1), np.arange(L + 1), chars, np.ones(L, np.int32), np.ones(L, np.int32))
    st.set_trees(w.father, w.left, w.right, w.age, w.root)
    alpha, beta, sbeta, ralpha = 3.0, 3000.0, 12000.0, 2.0
    Q = 3
    ta, tb = np.full(Q, alpha), np.array([sbeta, 1.0, beta])
    sm = gp.Sampler(st, w.pops, w.node_pop, theta_prior=(alpha, beta), tau_prior=(ta, tb), seed=11,
                    finetunes=(0.01, 0.6, 0.0008, 0.3), estimate_sample_age=[1, 0], locus_rate_finetune=0.8, rate_alpha=ralpha)
    sm.iterate(3000, trace=False)
    tr = sm.iterate(60000)
    assert sm.check()[0] == 0
    rng = np.random.default_rng(1)
    s0, t0 = rng.gamma(alpha, 1 / sbeta, 4_000_000), rng.gamma(alpha, 1 / beta, 4_000_000)
    keep = s0 < t0
    expect = {"theta_A": (alpha / beta, np.sqrt(alpha) / beta), "theta_root": (alpha / beta, np.sqrt(alpha) / beta),
              "tau_root": (t0[keep].mean(), t0[keep].std()), "sample_age_A": (s0[keep].mean(), s0[keep].std())}
    for col, name in [(0, "theta_A"), (2, "theta_root"), (3, "tau_root"), (4, "sample_age_A")]:
        x = tr[:, col]
        mean, sd = expect[name]
        se = batch_se(x)
        assert abs(x.mean() - mean) < 3.5 * se + 0.01 * mean, (name, x.mean(), mean, se)
        assert abs(x.std() - sd) < 0.12 * sd, (name, x.std(), sd)
    var = tr[:, 5] ** 2                                  # trace column = sqrt(mean (rate - 1)^2)
    want = (L - 1) / (ralpha * L + 1)
    assert abs(var.mean() - want) < 4.5 * batch_se(var) + 0.01 * want, (var.mean(), want)
    assert np.all(tr[:, -2] == 0.0)
    sm.close(); st.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/G-PhoCS-ref not built")
def test_posterior_means_match_the_reference_chain_with_ancient_samples():
    """BASELINE.json configs[4] shape (ancient samples in B, `locus-mut-rate VAR 1.0`) at 50 loci: posterior means of
    the thetas, taus and the rate spread against the reference's own chain on the same alignment, within 3 Monte-Carlo
    standard errors (+ 1 %); the estimated sample age, whose posterior is improper under the reference's prior, in order
    of magnitude (refchain.sample_age_columns)."""
    import refchain as rc
    L, iters = 50, 30000
    burn = iters // 5
    names, ref, model, w, ft, _ = rc.chain(rc.REF, "ref", "ancient", L, iters)
    Q, C = model.numPops, model.numCurPops
    K = 2 * Q - C
    assert names[1 + K].startswith("tau_") and names[2 + K] == "Variance-Mut", names
    ref = rc.parameter_columns(model, ref)[burn:]
    st = gp.LociStore.from_workload(w)
    st.set_rates(np.ones(w.L))
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=2024, finetunes=(ft["coal_time"], ft["theta"], ft["tau"], ft["mixing"]),
                    estimate_sample_age=estimated(model), locus_rate_finetune=0.3)
    tr = sm.iterate(iters)[burn:, :K + 2]
    assert sm.check()[0] == 0
    loose = rc.sample_age_columns(model)          # improper prior on the sample age: see refchain.sample_age_columns
    for k in range(K + 2):
        a, b = ref[:, k], tr[:, k]
        if k in loose:
            assert 0.25 < a.mean() / b.mean() < 4.0, (names[1 + k], a.mean(), b.mean())
            continue
        se = np.hypot(batch_se(a), batch_se(b))
        assert abs(a.mean() - b.mean()) < 3.0 * se + 0.01 * abs(a.mean()), (names[1 + k], a.mean(), b.mean(), se)
    sm.close(); st.close()


@pytest.mark.parametrize("cfg", ["ancient", "dip8mig"])
def test_trace_file_has_the_reference_format(cfg, tmp_path):
    """gphocsSamplerOpenTrace: header and row layout of performMCMC's trace file (GPhoCS.c:1255-1313, 1762-1769); where
    the reference binary is present its header for the same model is compared literally."""
    import subprocess
    model = synth.config(cfg)
    seq = str(tmp_path / "seqs.txt")
    w = synth.generate(model, 30, seed=7, seqfile=seq)
    st = gp.LociStore.from_workload(w)
    mig = (w.mig_start, w.mig_branch, w.mig_band, w.mig_age) if len(w.pops["band_src"]) else None
    extra = dict(estimate_sample_age=estimated(model), locus_rate_finetune=0.3) if cfg == "ancient" else {}
    sm = gp.Sampler(st, w.pops, w.node_pop, seed=3, migration=mig, **extra)
    path = str(tmp_path / "dev.trace")
    sm.open_trace(path, model.names, theta_tau_print=10000.0, mig_rate_print=0.001, sample_skip=1)
    rows = sm.iterate(6)
    sm.iterate(2, trace=False)
    sm.close_trace()
    names, tr = read_trace(path)
    assert list(tr[:, 0]) == [0, 2, 4, 6]
    K, B = 2 * sm.Q - sm.C, sm.B
    params = sm.width - 2
    factor = np.full(params, 10000.0)
    factor[K:K + B] = 0.001
    if cfg == "ancient":
        factor[-1] = 1.0
    for i, it in enumerate((0, 2, 4)):
        assert np.allclose(tr[i, 1:1 + params], rows[it, :params] * factor, rtol=0, atol=6e-6)
        assert abs(tr[i, 1 + params] - (rows[it, -2] + rows[it, -1]) / w.L) < 1e-6 and abs(tr[i, 2 + params] - rows[it, -2]) < 1e-6
    assert names[0] == "Sample" and names[-2:] == ["Data-ld-ln", "Full-ld-ln"] and len(names) == params + 3
    if os.path.exists(REF):
        ctl, rtrace = str(tmp_path / "ref.ctl"), str(tmp_path / "ref.trace")
        synth.write_control_file(model, ctl, seq, rtrace, iterations=20, seed=1, iterations_per_log=10)
        r = subprocess.run([REF, ctl], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout[-1500:]
        assert open(rtrace).readline() == open(path).readline()
    sm.close(); st.close()
