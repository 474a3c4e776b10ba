"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and the golden vectors
dumped from the reference.  Tolerance: north_star's 1e-10 relative on every per-locus log-likelihood
(observed ~1e-15: device exp/log are within 1 ulp of glibc's); genealogy statistics and log-density
are bit-exact; genealogy edits (trees, roots, flags) are exact."""
import importlib

import numpy as np
import pytest
import torch

import golden_io
import ops
from oracle import bindings as ob

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def oracle_loci(w, ids=None):
    out = []
    for l in (range(w.L) if ids is None else ids):
        p0, p1 = int(w.patt_start[l]), int(w.patt_start[l + 1])
        u0, u1 = int(w.unph_start[l]), int(w.unph_start[l + 1])
        lc = ob.OracleLocus(w.n, w.chars[p0:p1], w.num_phases[p0:p1], w.counts[u0:u1], float(w.rate[l]))
        lc.set_tree(w.father[l], w.left[l], w.right[l], w.age[l], int(w.root[l]))
        out.append(lc)
    return out


# ------------------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("name", golden_io.CASES)
def test_data_lnl_against_reference_golden(name):
    g = golden_io.load(name)
    st = gp.LociStore(int(g["n"]), g["patt_start"], g["unph_start"], g["chars"], g["num_phases"], g["counts"])
    st.set_trees(g["father"], g["left"], g["right"], g["age"], g["root"])
    st.set_rates(g["rate"])
    lnl, total = st.evaluate(0, want_sum=True)
    assert rel(lnl, g["data_lnl_full"]) < RTOL
    assert abs(total - g["data_lnl_full"].sum()) <= RTOL * abs(total)
    # conditional likelihood vectors of every internal node against the oracle
    n = int(g["n"])
    for l in range(0, int(g["L"]), 5):
        (p0, p1), (u0, u1), _, _ = golden_io.locus_slices(g, l)
        lc = ob.OracleLocus(n, g["chars"][p0:p1], g["num_phases"][p0:p1], g["counts"][u0:u1], float(g["rate"][l]))
        lc.set_tree(g["father"][l], g["left"][l], g["right"][l], g["age"][l], int(g["root"][l]))
        lc.compute(0)
        for node in range(2 * n - 1):
            got = st.clv(l, node, P=p1 - p0)
            assert np.allclose(got, lc.clv(node), rtol=1e-12, atol=0.0), (l, node)
    st.close()


@pytest.mark.parametrize("name", golden_io.CASES)
def test_genealogy_against_reference_golden(name):
    g = golden_io.load(name)
    pops = golden_io.pops_of(g)
    L, Q, B = int(g["L"]), int(g["Q"]), int(g["B"])
    gen = gp.Genealogy(L, pops)
    gen.set_events(g["ev_start"], g["pop_start"], g["ev_type"], g["ev_id"], g["ev_time"])
    r = gen.evaluate()
    pt, keep = ob.make_poptree(pops, g["band_start"], g["band_end"])
    lineages = gen.lineages()
    for l in range(L):
        _, _, (e0, e1), _ = golden_io.locus_slices(g, l)
        nl, cs, nc, ms, nm, lnl = ob.oracle_gen_locus(pt, g["pop_start"][l], g["ev_type"][e0:e1], g["ev_id"][e0:e1],
                                                      g["ev_time"][e0:e1])
        # same elapsed times, same order of operations, no FMA: bit-exact against the oracle
        assert np.array_equal(r["coal"][l], cs)
        assert np.array_equal(r["num_coals"][l], nc)
        assert np.array_equal(r["mig"][l], ms)
        assert np.array_equal(r["num_migs"][l], nm)
        assert r["lnl"][l] == lnl
        assert np.array_equal(lineages[e0:e1], nl)
        assert np.array_equal(lineages[e0:e1], g["ev_lineages"][e0:e1])
    # against the reference's own (incrementally maintained) numbers
    assert rel(r["lnl"], g["gen_lnl"]) < RTOL
    assert np.allclose(r["coal"], g["coal_stats"], rtol=RTOL, atol=1e-13)
    assert np.array_equal(r["num_coals"], g["num_coals"])
    assert np.array_equal(r["num_migs"], g["num_migs"])
    # computeTotalStats
    assert np.array_equal(r["total_num_coals"], g["total_num_coals"])
    assert np.array_equal(r["total_num_migs"], g["total_num_migs"])
    assert np.allclose(r["total_coal"], g["total_coal_stats"], rtol=1e-9)
    assert np.allclose(r["total_mig"], g["total_mig_stats"], rtol=1e-9, atol=1e-12)
    assert abs(r["sum_lnl"] - r["lnl"].sum()) <= 1e-12 * abs(r["sum_lnl"])
    gen.close()


# ------------------------------------------------------------------------------------- scalar (drop-in) API
def test_scalar_api_replays_reference_traces():
    """The reference's LocusData entry points, exported by the CUDA library, replay the proposal traces
    recorded from the reference (adjust age / SPR / rescale / ancient-sample / rate moves, accept & reject)."""
    z = np.load(golden_io.os.path.join(golden_io.HERE, "golden", "ops_reference.npz"))
    cases = []
    for i in range(int(z["num_cases"])):
        cases.append({k[len(f"c{i}_"):]: z[k] for k in z.files if k.startswith(f"c{i}_")})
    # all loci of one process must share the leaf count: group traces by n
    by_n = {}
    for c in cases:
        by_n.setdefault(int(c["n"]), []).append(c)
    for n, group in by_n.items():
        loci = []
        for c in group:
            lc = gp.ScalarLocus(n, c["chars"], c["num_phases"], c["counts"], float(c["rate"]))
            lc.set_tree(c["father"], c["left"], c["right"], c["age"], int(c["root"]))
            loci.append(lc)
        for c, lc in zip(group, loci):
            tr = ops.run_ops(lc, n, int(c["seed"]), int(c["steps"]), allow_leaf_age=bool(c["leaf"]), rate_moves=bool(c["ratem"]))
            lnls, trees = [], []
            for e in tr:
                if e[0] in ("init", "final-full", "rate"):
                    lnls.append([e[1], np.nan])
                else:
                    lnls.append([e[1], e[3]])
                    trees.append(np.concatenate([e[4], e[5], e[6], [e[8]]]).astype(np.float64).tolist() + e[7].tolist())
            lnls = np.array(lnls)
            ok = ~np.isnan(c["lnls"])
            assert rel(lnls[ok], c["lnls"][ok]) < RTOL
            assert np.array_equal(np.array(trees), c["trees"])     # genealogy edits are exact
            assert lc.check() == 1
        store_mismatch = gp.lib().gphocsStoreCheckMirror(gp.lib().gpuLociStore())
        assert store_mismatch == 0
        for lc in loci:
            lc.free()


# ------------------------------------------------------------------------------------- batched engine
@pytest.mark.parametrize("cfg,L", [("hap16", 300), ("dip8mig", 200), ("pop6mig4", 150), ("ancient", 200)])
def test_batched_proposals_against_oracle(cfg, L):
    """One proposal per locus per round for all loci at once (the loop-interchanged Shape A / Shape B
    traffic of GPhoCS.c:2287-4916), random accept/reject masks, against per-locus oracle instances."""
    w = synth.generate(synth.config(cfg), L, seed=11, missing_frac=0.05 if cfg == "dip8mig" else 0.0)
    st = gp.LociStore.from_workload(w)
    st.set_debug(True)
    orc = oracle_loci(w)
    want = np.array([o.compute(0) for o in orc])
    got = st.evaluate(0)
    assert rel(got, want) < RTOL
    for o in orc:
        o.reset()
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    rng = np.random.default_rng(3)
    n = w.n

    class Rec:
        """records the edits a proposal makes so they can be sent as one batch"""
        def __init__(self, o, l):
            self.o, self.l, self.ops = o, l, []
        def tree(self): return self.o.tree()
        def adjust_age(self, node, age):
            self.ops.append((self.l, gp.OP_ADJUST_AGE, node, 0, age)); return self.o.adjust_age(node, age)
        def spr(self, s, t, age):
            self.ops.append((self.l, gp.OP_SPR, s, t, age)); return self.o.spr(s, t, age)
        def scale_all(self, f):
            self.ops.append((self.l, gp.OP_SCALE_ALL, 0, 0, f)); return self.o.scale_all(f)

    for rnd in range(12):
        recs, spr_status = [], {}
        for l in range(L):
            r = Rec(orc[l], l)
            d = ops.propose(r, rng, n, allow_leaf_age=(cfg == "ancient"))
            if d[0] == "spr":
                spr_status[l] = d[4]
            recs += r.ops
        batch = np.array(recs, dtype=gp.OP_DTYPE)
        status = st.apply_ops(batch, want_status=True)
        for i, rec in enumerate(recs):
            if rec[1] == gp.OP_SPR:
                assert status[i] == spr_status[rec[0]]
        got, total = st.evaluate(1, want_sum=True)
        want = np.array([o.compute(1) if not any(r[0] == l and r[1] == gp.OP_SCALE_ALL for r in recs) else o.lnl()
                         for l, o in enumerate(orc)]) if False else None
        scaled = {r[0] for r in recs if r[1] == gp.OP_SCALE_ALL}
        want = np.array([o.lnl() if l in scaled else o.compute(1) for l, o in enumerate(orc)])
        assert rel(got, want) < RTOL, rnd
        assert abs(total - want.sum()) <= RTOL * abs(total)
        accept = rng.random(L) < 0.5
        for l, o in enumerate(orc):
            o.reset() if accept[l] else o.revert()
        st.apply_ops(gp.make_ops(np.arange(L), np.where(accept, gp.OP_COMMIT, gp.OP_REVERT)))
        f, lf, rt, a, root = st.get_trees()
        for l, o in enumerate(orc):
            of, ol, orr, oa, oroot = o.tree()
            assert np.array_equal(f[l], of) and np.array_equal(lf[l], ol) and np.array_equal(rt[l], orr)
            assert np.array_equal(a[l], oa) and root[l] == oroot
        assert rel(st.lnl(), [o.lnl() for o in orc]) < RTOL
        assert st.check_mirror() == 0
    # the incrementally maintained state equals a from-scratch evaluation (the reference's checkAll invariant)
    inc = st.lnl()
    full = st.evaluate(0)
    assert rel(full, inc) < 1e-9
    assert rel(full, [o.compute(0) for o in orc]) < RTOL
    st.close()


def test_oversized_and_ragged_loci():
    """Loci wider than one CTA (P > 128), single-column loci, loci without live patterns, missing data."""
    rng = np.random.default_rng(8)
    n = 10
    Ps = [300, 1, 0, 129, 128, 2, 517, 0, 7]
    chars, phases, counts, ps, us = [], [], [], [0], [0]
    for P in Ps:
        c, ph, ct = ops.random_patterns(n, P, rng, diploid_pairs=5, missing=0.1) if P else (
            np.zeros((0, n), np.uint8), np.zeros(0, np.int32), np.zeros(0, np.int32))
        chars.append(c); phases.append(ph); counts.append(ct)
        ps.append(ps[-1] + len(ph)); us.append(us[-1] + len(ct))
    st = gp.LociStore(n, ps, us, np.concatenate(chars), np.concatenate(phases), np.concatenate(counts))
    trees = [ops.random_tree(n, rng) for _ in Ps]
    st.set_trees(*[np.array([t[k] for t in trees]) for k in range(4)], np.array([t[4] for t in trees]))
    got, total = st.evaluate(0, want_sum=True)
    want = []
    for i, P in enumerate(Ps):
        lc = ob.OracleLocus(n, chars[i], phases[i], counts[i], 1.0)
        lc.set_tree(*trees[i])
        want.append(lc.compute(0))
    assert rel(got, want) < RTOL
    assert got[2] == 0.0 and got[7] == 0.0
    assert abs(total - sum(want)) <= RTOL * abs(total)
    st.close()


def test_subset_evaluation_touches_only_listed_loci():
    w = synth.generate(synth.config("hap16"), 64, seed=2)
    st = gp.LociStore.from_workload(w)
    full = st.evaluate(0).copy()
    st.apply_ops(gp.make_ops(np.arange(64), gp.OP_COMMIT))
    ids = np.array([3, 17, 40], np.int32)
    node = w.n + 2
    st.apply_ops(gp.make_ops(np.arange(64), gp.OP_ADJUST_AGE, a=node, x=w.age[:, node] * 1.01))
    part = st.evaluate(1, ids=ids)
    now = st.lnl()
    untouched = np.setdiff1d(np.arange(64), ids)
    assert np.array_equal(now[untouched], full[untouched])
    assert np.all(now[ids] != full[ids]) and np.array_equal(part, now[ids])
    st.close()


def test_rejects_bad_input():
    with pytest.raises(RuntimeError):
        gp.LociStore(3, [0, 1], [0, 1], np.frombuffer(b"TXT", np.uint8).reshape(1, 3), [1], [5])     # bad character
    with pytest.raises(RuntimeError):
        gp.LociStore(3, [0, 2], [0, 1], np.frombuffer(b"TCTTCA", np.uint8).reshape(2, 3), [4, 0], [5])  # phase group overruns


# ------------------------------------------------------------------------------------- BASELINE sizes: properties
@pytest.mark.parametrize("cfg", ["hap16", "dip8mig"])
def test_full_size_properties(cfg):
    """10k loci (BASELINE.json configs[1], [2]): size-independent properties — a sampled oracle check,
    revert restores bit-identical likelihoods, rescaling by f then 1/f returns to the start, the incremental
    path equals the full path, and the device-side sum equals the sum of the per-locus values."""
    L = synth.CONFIG_LOCI[cfg]
    w = synth.generate(synth.config(cfg), L, seed=1000)
    st = gp.LociStore.from_workload(w)
    base, total = st.evaluate(0, want_sum=True)
    assert abs(total - base.sum()) <= 1e-12 * abs(total)
    ids = np.random.default_rng(0).choice(L, 200, replace=False)
    orc = oracle_loci(w, ids)
    assert rel(base[ids], [o.compute(0) for o in orc]) < RTOL
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    # move one internal node in every locus, evaluate incrementally, reject: bit-identical restore
    node = w.n + 1
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_ADJUST_AGE, a=node, x=w.age[:, node] * (1 + 1e-3)))
    moved = st.evaluate(1)
    assert np.all(moved != base)
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_REVERT))
    assert np.array_equal(st.lnl(), base)
    assert np.array_equal(st.evaluate(1), base)          # nothing dirty: stored values come back
    # accept the same move: incremental == from scratch
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_ADJUST_AGE, a=node, x=w.age[:, node] * (1 + 1e-3)))
    inc = st.evaluate(1).copy()
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    assert rel(st.evaluate(0), inc) < 1e-12
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    # mixing-style rescale (scaleAllNodeAges on every locus) and its inverse
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_SCALE_ALL, x=1.25))
    st.evaluate(1)
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_SCALE_ALL, x=0.8))
    back = st.evaluate(1)
    assert rel(back, inc) < 1e-9
    st.close()


def test_full_size_genealogy_properties():
    """100k-locus shape of configs[3] at 20k loci: totals equal the sums of per-locus statistics, counts
    are exact integers, theta-rescaling moves lnL by the closed form UpdateTheta uses (GPhoCS.c:3068-3070)."""
    m = synth.config("pop6mig4")
    L = 20000
    w = synth.generate(m, L, seed=77)
    gen = gp.Genealogy(L, w.pops)
    gen.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    r = gen.evaluate()
    assert np.array_equal(r["total_num_coals"], r["num_coals"].sum(0))
    assert r["total_num_coals"].sum() == L * (w.n - 1)
    assert np.array_equal(r["total_num_migs"], r["num_migs"].sum(0))
    assert r["total_num_migs"].sum() == len(w.mig_age)
    assert np.allclose(r["total_coal"], r["coal"].sum(0), rtol=1e-11)
    assert np.allclose(r["total_mig"], r["mig"].sum(0), rtol=1e-11)
    assert abs(r["sum_lnl"] - r["lnl"].sum()) <= 1e-11 * abs(r["sum_lnl"])
    # oracle on a sample
    pt, keep = ob.make_poptree(w.pops, w.band_start, w.band_end)
    for l in np.random.default_rng(1).choice(L, 100, replace=False):
        e0, e1 = int(w.ev_start[l]), int(w.ev_start[l + 1])
        nl, cs, nc, ms, nm, lnl = ob.oracle_gen_locus(pt, w.pop_start[l], w.ev_type[e0:e1], w.ev_id[e0:e1], w.ev_time[e0:e1])
        assert r["lnl"][l] == lnl and np.array_equal(r["coal"][l], cs) and np.array_equal(r["mig"][l], ms)
    # UpdateTheta's closed form
    theta2 = w.pops["theta"].copy()
    theta2[3] *= 1.1
    gen.set_params(theta2, w.pops["band_rate"])
    r2 = gen.evaluate(per_locus_stats=False)
    lnc = np.log(1.1)
    delta = -(lnc * r["total_num_coals"][3] + (1 / theta2[3] - 1 / w.pops["theta"][3]) * r["total_coal"][3])
    assert abs((r2["sum_lnl"] - r["sum_lnl"]) - delta) <= 1e-9 * abs(delta)
    gen.close()


def test_page_locked_genealogies_take_the_direct_route():
    """gphocsStoreSetTrees with page-locked arrays covering all loci copies them as they are (no staging); mirror,
    device state and likelihoods must equal those of the staged route."""
    w = synth.generate(synth.config("dip8mig"), 3000, seed=21)
    a = gp.LociStore.from_workload(w)
    b = gp.LociStore(w.n, w.patt_start, w.unph_start, w.chars, w.num_phases, w.counts)
    hw = {k: gp.pinned_like(getattr(w, k)) for k in ("father", "left", "right", "age", "root")}
    k0 = gp.lib().gphocsKernelLaunchCount()
    b.set_trees(hw["father"], hw["left"], hw["right"], hw["age"], hw["root"])
    assert gp.lib().gphocsKernelLaunchCount() == k0 + 1
    b.set_rates(w.rate)
    for x, y in zip(a.get_trees(), b.get_trees()):
        assert np.array_equal(x, y)
    assert a.check_mirror() == 0 and b.check_mirror() == 0
    assert np.array_equal(a.evaluate(0), b.evaluate(0))
    # a second full set after proposals were made and rejected: buffer selectors stay consistent on both sides
    ops = gp.make_ops(np.arange(w.L), gp.OP_ADJUST_AGE, a=w.n + 1, x=w.age[:, w.n + 1] * 1.001)
    for st in (a, b):
        st.apply_ops(gp.make_ops(np.arange(w.L), gp.OP_COMMIT)); st.apply_ops(ops); st.evaluate(1)
        st.apply_ops(gp.make_ops(np.arange(w.L), gp.OP_REVERT))
    a.set_trees(w.father, w.left, w.right, w.age, w.root)
    b.set_trees(hw["father"], hw["left"], hw["right"], hw["age"], hw["root"])
    assert b.check_mirror() == 0
    assert np.array_equal(a.evaluate(0), b.evaluate(0))
    # the direct route leaves the host mirror to whoever reads it next: an edit batch right after it must land on
    # the new genealogies on both sides
    b.set_trees(hw["father"], hw["left"], hw["right"], hw["age"], hw["root"])
    for st in (a, b):
        st.apply_ops(ops)
    for x, y in zip(a.get_trees(), b.get_trees()):
        assert np.array_equal(x, y)
    assert a.check_mirror() == 0 and b.check_mirror() == 0
    assert np.array_equal(a.evaluate(1), b.evaluate(1))
    a.close(); b.close()


def test_page_locked_event_snapshot_takes_the_direct_route():
    """gphocsGenSetEvents with page-locked arrays: copied as they are and narrowed on the device; same statistics and
    log-densities as the staged route, and a malformed snapshot is still refused."""
    w = synth.generate(synth.config("pop6mig4"), 2500, seed=22)
    res = []
    for pinned in (False, True):
        gen = gp.Genealogy(w.L, w.pops)
        arrs = [getattr(w, k) for k in ("ev_start", "pop_start", "ev_type", "ev_id", "ev_time")]
        if pinned:
            arrs = [gp.pinned_like(a) for a in arrs]
        gen.set_events(*arrs)
        res.append(gen.evaluate())
        if pinned:
            bad = [a.copy() for a in arrs]
            bad = [gp.pinned_like(a) for a in bad]
            bad[2][5] = 99                                   # not an event type
            with pytest.raises(RuntimeError):
                gen.set_events(*bad)
        gen.close()
    a, b = res
    for key in ("lnl", "coal", "num_coals", "mig", "num_migs", "total_coal", "total_num_coals", "total_mig", "total_num_migs"):
        assert np.array_equal(a[key], b[key]), key
    assert a["sum_lnl"] == b["sum_lnl"]


def test_packed_wire_formats_equal_the_plain_ones():
    """gphocsStoreSetTreesPacked / gphocsGenSetEventsPacked (16-bit topology and event codes on the wire) leave the
    device in the state gphocsStoreSetTrees / gphocsGenSetEvents leave it in: genealogies, mirror, every
    log-likelihood and statistic bit for bit; malformed packed input is refused."""
    w = synth.generate(synth.config("pop6mig4"), 2000, seed=23)
    a = gp.LociStore.from_workload(w)
    b = gp.LociStore(w.n, w.patt_start, w.unph_start, w.chars, w.num_phases, w.counts)
    topo = gp.pack_trees(w.father, w.left, w.right)
    assert topo.dtype == np.int16 and topo.shape == (w.L, 2 * w.n - 1, 3)
    for pinned in (False, True):
        t, ag, ro = (gp.pinned_like(x) for x in (topo, w.age, w.root)) if pinned else (topo, w.age, w.root)
        b.set_trees_packed(t, ag, ro)
        if not pinned:
            b.set_rates(w.rate)
        for x, y in zip(a.get_trees(), b.get_trees()):
            assert np.array_equal(x, y)
        assert b.check_mirror() == 0
        assert np.array_equal(a.evaluate(0), b.evaluate(0))
    bad = topo.copy()
    bad[7, 3, 1] = 2 * w.n - 1                                # a son outside the tree
    with pytest.raises(RuntimeError):
        b.set_trees_packed(bad, w.age, w.root)
    a.close(); b.close()

    es, ps, code = gp.pack_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id)
    assert es.dtype == np.int32 and ps.dtype == np.uint16 and code.dtype == np.uint16
    ga = gp.Genealogy(w.L, w.pops)
    ga.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    ra = ga.evaluate()
    gb = gp.Genealogy(w.L, w.pops)
    for pinned in (False, True):
        arrs = (es, ps, code, w.ev_time)
        if pinned:
            arrs = tuple(gp.pinned_like(x) for x in arrs)
        gb.set_events_packed(*arrs)
        rb = gb.evaluate()
        for key in ("lnl", "coal", "num_coals", "mig", "num_migs", "total_coal", "total_num_coals", "total_mig", "total_num_migs"):
            assert np.array_equal(ra[key], rb[key]), key
        assert ra["sum_lnl"] == rb["sum_lnl"]
    nb = len(w.pops["band_src"])
    for k, v in ((5, 7 | (1 << 3)),                           # a band id on an event that carries none
                 (int(np.flatnonzero((code & 7) == 3)[0]), 3 | (nb << 3))):   # a band that does not exist
        wrong = code.copy()
        wrong[k] = v
        with pytest.raises(RuntimeError):
            gb.set_events_packed(es, ps, wrong, w.ev_time)
    wrong = ps.copy()
    wrong[11, 2] = wrong[11, 3] + 1                           # chains out of order
    with pytest.raises(RuntimeError):
        gb.set_events_packed(es, wrong, code, w.ev_time)
    ga.close(); gb.close()


# ------------------------------------------------------------------------------------- edge lengths, full sizes
def test_zero_tiny_and_negative_edge_lengths_through_the_kernel():
    """computeEdgeConditionalJC (LocusDataLikelihood.c:1843-1845): an edge shorter than 1e-100 — zero (a child as old as
    its father), denormal-small or NEGATIVE (a child older than its father, which proposals pass through before they
    are rejected) — has transition probability exactly 0, i.e. the child's vector is copied.  Full and incremental
    evaluation of such genealogies against the oracle, internal vectors included."""
    w = synth.generate(synth.config("dip8mig"), 96, seed=23)
    n, N = w.n, 2 * w.n - 1
    age = w.age.copy()
    kinds = []
    for l in range(w.L):
        v = n + (l % (n - 2))                      # an internal node that is not the root in most loci
        f = int(w.father[l, v])
        if f < 0:
            v = int(w.left[l, v]) if int(w.left[l, v]) >= n else int(w.right[l, v])
            f = int(w.father[l, v])
        if f < 0 or v < n:
            kinds.append(None)
            continue
        k = l % 4
        if k == 0:
            age[l, v] = age[l, f]                  # zero-length edge
        elif k == 1:
            age[l, v] = age[l, f] * (1.0 + 1e-3)   # negative edge length: t < 1e-100 holds as well
        elif k == 2:
            age[l, v] = np.nextafter(age[l, f], 0.0)    # one ulp below the father: ~1e-19 > 1e-100, NOT the branch
        else:
            age[l, f] = age[l, v]                  # father pulled down onto the child (both its edges change)
        kinds.append((v, f, k))
    w2 = w
    st = gp.LociStore.from_workload(w2)
    st.set_trees(w.father, w.left, w.right, age, w.root)
    got = st.evaluate(0)
    orc = oracle_loci(w)
    for l, o in enumerate(orc):
        o.set_tree(w.father[l], w.left[l], w.right[l], age[l], int(w.root[l]))
    want = np.array([o.compute(0) for o in orc])
    finite = np.isfinite(want)
    assert finite.sum() > 0.9 * w.L
    assert rel(got[finite], want[finite]) < RTOL
    assert np.array_equal(np.isfinite(got), finite)            # a likelihood the reference loses, this path loses too
    for l in range(0, w.L, 7):
        P = int(w.patt_start[l + 1] - w.patt_start[l])
        for node in range(n, N):
            assert np.allclose(st.clv(l, node, P=P), orc[l].clv(node), rtol=1e-12, atol=0.0), (l, node, kinds[l])
    # the same states reached by an incremental evaluation: back to the original ages, then the edit as a proposal
    st.apply_ops(gp.make_ops(np.arange(w.L), gp.OP_COMMIT))
    st.set_trees(w.father, w.left, w.right, w.age, w.root)
    st.evaluate(0)
    st.apply_ops(gp.make_ops(np.arange(w.L), gp.OP_COMMIT))
    sel = [l for l, kd in enumerate(kinds) if kd is not None and kd[2] != 3]
    nodes = np.array([kinds[l][0] for l in sel])
    st.apply_ops(gp.make_ops(np.array(sel), gp.OP_ADJUST_AGE, a=nodes, x=age[sel, nodes]))
    inc = st.evaluate(1)
    assert rel(inc[sel][finite[sel]], want[sel][finite[sel]]) < RTOL
    st.close()


@pytest.mark.parametrize("cfg,L", [("pop6mig4", 100_000), ("ancient", 50_000)])
def test_full_size_data_and_genealogy_likelihood_against_sampled_oracle(cfg, L):
    """BASELINE.json configs[3] (100k loci x 24 leaves, 4 bands) and configs[4] (50k loci, ancient samples, per-locus
    rates) at FULL size: every locus is evaluated on the device; a random sample of 600 loci is checked against the
    oracle (data lnL to 1e-10, genealogy statistics and lnL bit for bit), and the device totals against the sum of
    the per-locus values."""
    w = synth.generate(synth.config(cfg), L, seed=2026)
    st = gp.LociStore.from_workload(w)
    lnl, total = st.evaluate(0, want_sum=True)
    assert np.all(np.isfinite(lnl))
    rng = np.random.default_rng(5)
    ids = np.sort(rng.choice(L, 600, replace=False))
    want = np.array([o.compute(0) for o in oracle_loci(w, ids)])
    assert rel(lnl[ids], want) < RTOL
    assert abs(total - lnl.sum()) <= 1e-11 * abs(total)
    gen = gp.Genealogy(L, w.pops)
    gen.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    r = gen.evaluate()
    pt, keep = ob.make_poptree(w.pops, w.band_start, w.band_end)
    B = len(w.pops["band_src"])
    for l in ids:
        e0, e1 = int(w.ev_start[l]), int(w.ev_start[l + 1])
        _, cs, nc, ms, nm, glnl = ob.oracle_gen_locus(pt, w.pop_start[l], w.ev_type[e0:e1], w.ev_id[e0:e1], w.ev_time[e0:e1])
        assert r["lnl"][l] == glnl and np.array_equal(r["coal"][l], cs) and np.array_equal(r["num_coals"][l], nc)
        assert np.array_equal(r["mig"][l][:B], ms[:B]) and np.array_equal(r["num_migs"][l][:B], nm[:B])
    assert np.array_equal(r["total_num_coals"], r["num_coals"].sum(0))
    assert abs(r["sum_lnl"] - r["lnl"].sum()) <= 1e-11 * abs(r["sum_lnl"])
    st.close(); gen.close()


@pytest.mark.parametrize("cfg", ["pop6mig4", "dip8mig", "hap16"])
def test_incremental_recalc_equals_full_evaluation(cfg):
    """gphocsGenRecalc = recalcStats (patch.c:2387-2513) for single (locus, population) chains whose elapsed times
    changed, as after rubberBand: the stored statistics become bit for bit those of a full evaluation of the updated
    snapshot, the returned value is recalcStats' (the same operations in the same order), and sending the old times
    back restores the old statistics exactly (what a rejected UpdateTau proposal does)."""
    w = synth.generate(synth.config(cfg), 3000, seed=77)
    Q, B = len(w.pops["father"]), len(w.pops["band_src"])
    theta = np.full(Q, 1e-3) * (1 + 0.1 * np.arange(Q))
    rate = np.array([150.0 + 30 * b for b in range(B)])
    gen = gp.Genealogy(w.L, w.pops)
    gen.set_params(theta, rate)
    gen.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    before = gen.evaluate()
    rng = np.random.default_rng(11)
    loci = np.sort(rng.choice(w.L, 1200, replace=False))
    pops = rng.integers(0, Q, len(loci))
    new_time = w.ev_time.copy()
    starts, chunks = [0], []
    for l, p in zip(loci, pops):
        a = int(w.ev_start[l] + w.pop_start[l][p]); b = int(w.ev_start[l] + w.pop_start[l][p + 1])
        f = rng.uniform(0.6, 1.4, b - a)                       # every interval of the chain rescaled on its own
        new_time[a:b] = w.ev_time[a:b] * f
        chunks.append(new_time[a:b].copy())
        starts.append(starts[-1] + (b - a))
    delta = gen.recalc(loci, pops, starts, np.concatenate(chunks) if chunks else np.zeros(0))
    after_inc = gen.stats_only()
    # the same snapshot evaluated from scratch
    gen2 = gp.Genealogy(w.L, w.pops)
    gen2.set_params(theta, rate)
    gen2.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, new_time)
    full = gen2.evaluate()
    assert np.array_equal(after_inc["coal"], full["coal"]) and np.array_equal(after_inc["mig"], full["mig"])
    # recalcStats' return value from the stored statistics before and after, same operations in the same order
    band_tgt = np.asarray(w.pops["band_tgt"])
    for k, (l, p) in enumerate(zip(loci, pops)):
        d = 0.0
        a = int(w.ev_start[l] + w.pop_start[l][p]); b = int(w.ev_start[l] + w.pop_start[l][p + 1])
        for e in range(a, b):                                  # MIG_BAND_END events in chain order
            if w.ev_type[e] == 4:
                bid = int(w.ev_id[e])
                d = d - (full["mig"][l][bid] - before["mig"][l][bid]) * rate[bid]
        d = d - (full["coal"][l][p] - before["coal"][l][p]) / theta[p]
        assert delta[k] == d, (k, l, p, delta[k], d)
        assert all(band_tgt[int(w.ev_id[e])] == p for e in range(a, b) if w.ev_type[e] == 4)
    # the rejected proposal: old times back
    old_chunks = [w.ev_time[int(w.ev_start[l] + w.pop_start[l][p]):int(w.ev_start[l] + w.pop_start[l][p + 1])] for l, p in zip(loci, pops)]
    back = gen.recalc(loci, pops, starts, np.concatenate(old_chunks))
    restored = gen.stats_only()
    assert np.array_equal(restored["coal"], before["coal"]) and np.array_equal(restored["mig"], before["mig"])
    assert np.allclose(back, -delta, rtol=1e-9, atol=1e-9)
    # a chain of the wrong length is refused and nothing changes
    with pytest.raises(RuntimeError):
        gen.recalc(loci[:1], pops[:1], [0, starts[1] + 1], np.ones(starts[1] + 1))
    assert np.array_equal(gen.stats_only()["coal"], before["coal"])
    gen.close(); gen2.close()


def test_delta_upload_step_equals_the_full_upload():
    """The delta routes of an MCMC-shaped host step — gphocsStoreApplyOpsAsync (edit records for the device copy only,
    no waiting) and gphocsGenRecalcAsync (new elapsed times of single chains) — leave the device in exactly the state
    a full upload of the edited genealogies and event snapshot leaves it in: log-likelihoods and statistics bit for bit;
    the host mirror catches up on demand, pending proposal included (revert through the synchronous API)."""
    w = synth.generate(synth.config("dip8mig"), 2000, seed=5)
    L, n = w.L, w.n
    st = gp.LociStore.from_workload(w)
    st.evaluate(0)
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
    node = n + 3
    new_age = w.age[:, node] * 1.0005
    recs = np.zeros(2 * L, gp.OP_DTYPE)                       # per locus: resetSaved, then adjustGenNodeAge
    recs["locus"] = np.repeat(np.arange(L), 2)
    recs["type"] = np.tile([gp.OP_COMMIT, gp.OP_ADJUST_AGE], L)
    recs["a"] = np.tile([0, node], L)
    recs["x"][1::2] = new_age
    st.apply_ops_async(gp.pinned_like(recs))
    got = st.evaluate(0)
    age2 = w.age.copy(); age2[:, node] = new_age
    st2 = gp.LociStore.from_workload(w)
    st2.set_trees(w.father, w.left, w.right, age2, w.root)
    want = st2.evaluate(0)
    assert np.array_equal(got, want)
    # the mirror follows: trees, the pending proposal's saved copies, then a revert through the synchronous API
    f, l_, r, a, root = st.get_trees()
    assert np.array_equal(a, age2)
    assert st.check_mirror() == 0
    st.apply_ops(gp.make_ops(np.arange(L), gp.OP_REVERT))
    assert np.array_equal(st.get_trees()[3], w.age) and st.check_mirror() == 0
    # records out of order or outside the tree are refused on the device and reported by the next sync
    bad = np.zeros(2, gp.OP_DTYPE); bad["locus"] = [5, 3]; bad["type"] = gp.OP_COMMIT
    st.apply_ops_async(gp.pinned_like(bad))
    with pytest.raises(RuntimeError):
        st.sync()
    st.sync()
    # genealogy: new times of one chain per locus, asynchronously, against a full snapshot
    Q = len(w.pops["father"])
    gen = gp.Genealogy(L, w.pops)
    gen.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, w.ev_time)
    gen.evaluate()
    pops = (np.arange(L) % Q).astype(np.int32)
    new_time = w.ev_time.copy()
    starts, chunks = [0], []
    for l in range(L):
        a0 = int(w.ev_start[l] + w.pop_start[l][pops[l]]); b0 = int(w.ev_start[l] + w.pop_start[l][pops[l] + 1])
        new_time[a0:b0] *= 0.9
        chunks.append(new_time[a0:b0]); starts.append(starts[-1] + b0 - a0)
    pin = [gp.pinned_like(np.arange(L, dtype=np.int32)), gp.pinned_like(pops), gp.pinned_like(np.asarray(starts, np.int32)),
           gp.pinned_like(np.concatenate(chunks))]
    assert gen.recalc_async(*pin)
    gen.sync()
    full = gp.Genealogy(L, w.pops)
    full.set_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id, new_time)
    ref = full.evaluate()
    inc = gen.evaluate()                                       # full evaluation of the snapshot the deltas produced
    assert np.array_equal(inc["lnl"], ref["lnl"]) and np.array_equal(inc["coal"], ref["coal"]) and np.array_equal(inc["mig"], ref["mig"])
    st.close(); st2.close(); gen.close(); full.close()
