"""CPU: the oracle restatement against the golden vectors dumped from the reference.
Same CPU, same operation order, no FMA contraction => bit-exact (== on doubles)."""
import ctypes as C

import numpy as np
import pytest

import golden_io
import ops
from oracle import bindings as ob


@pytest.mark.parametrize("name", golden_io.CASES)
def test_data_lnl_matches_reference(name):
    g = golden_io.load(name)
    n = int(g["n"])
    for l in range(int(g["L"])):
        (p0, p1), (u0, u1), _, _ = golden_io.locus_slices(g, l)
        lc = ob.OracleLocus(n, g["chars"][p0:p1], g["num_phases"][p0:p1], g["counts"][u0:u1], float(g["rate"][l]))
        lc.set_tree(g["father"][l], g["left"][l], g["right"][l], g["age"][l], int(g["root"][l]))
        v = lc.compute(0)
        assert v == g["data_lnl_full"][l]
        # the reference's incrementally maintained value agrees with its own full recompute to 1e-9 rel
        # (checkLocusDataLikelihood, LocusDataLikelihood.c:724); so must we
        assert abs(v - g["data_lnl"][l]) <= 1e-9 * abs(v)
    assert abs(g["data_lnl"].sum() - float(g["total_data_lnl"])) <= 1e-9 * abs(g["data_lnl"].sum())


@pytest.mark.parametrize("name", golden_io.CASES)
def test_genealogy_stats_and_lnl_match_reference(name):
    g = golden_io.load(name)
    pt, keep = ob.make_poptree(golden_io.pops_of(g), g["band_start"], g["band_end"])
    Q, B = int(g["Q"]), int(g["B"])
    tot_c, tot_m = np.zeros(Q), np.zeros(B)
    for l in range(int(g["L"])):
        _, _, (e0, e1), _ = golden_io.locus_slices(g, l)
        nl, cs, nc, ms, nm, lnl = ob.oracle_gen_locus(pt, g["pop_start"][l], g["ev_type"][e0:e1], g["ev_id"][e0:e1],
                                                      g["ev_time"][e0:e1])
        assert np.array_equal(nl, g["ev_lineages"][e0:e1])
        assert np.array_equal(nc, g["num_coals"][l])
        assert np.array_equal(nm, g["num_migs"][l])
        # the reference maintains its statistics incrementally (considerEventMove deltas); a from-scratch
        # recomputation agrees to checkGtreeStructure's 1e-10 (patch.c:2986), not bit-for-bit
        assert np.allclose(cs, g["coal_stats"][l], rtol=1e-10, atol=1e-13)
        assert np.allclose(ms, g["mig_stats"][l], rtol=1e-10, atol=1e-13)
        ref_lnl = float(g["gen_lnl"][l])
        assert abs(lnl - ref_lnl) <= 1e-10 * abs(ref_lnl)
        # gtreeLnLikelihood on the reference's own stored statistics: bit-exact
        lib = ob.oracle()
        exact = lib.orc_gen_lnl(C.byref(pt), ob.dp(np.ascontiguousarray(g["coal_stats"][l])),
                                ob.ip(np.ascontiguousarray(g["num_coals"][l])),
                                ob.dp(np.ascontiguousarray(np.resize(g["mig_stats"][l], max(B, 1)))),
                                ob.ip(np.ascontiguousarray(np.resize(g["num_migs"][l], max(B, 1)).astype(np.int32))))
        assert exact == ref_lnl
        tot_c += cs
        tot_m += ms
    assert np.allclose(tot_c, g["total_coal_stats"], rtol=1e-9)
    assert np.allclose(tot_m, g["total_mig_stats"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(g["num_coals"].sum(0), g["total_num_coals"])


@pytest.mark.parametrize("name", golden_io.CASES)
def test_event_construction_matches_reference_chains(name):
    """constructEventChain restated: same event order per population and elapsed times to 1e-12
    (the reference's chains have been edited incrementally for a few MCMC iterations)."""
    g = golden_io.load(name)
    pt, keep = ob.make_poptree(golden_io.pops_of(g), g["band_start"], g["band_end"])
    lib = ob.oracle()
    n, Q = int(g["n"]), int(g["Q"])
    for l in range(int(g["L"])):
        _, _, (e0, e1), (m0, m1) = golden_io.locus_slices(g, l)
        E = e1 - e0
        ps = np.zeros(Q + 1, np.int32)
        ty, idd = np.zeros(E + 8, np.int32), np.zeros(E + 8, np.int32)
        el = np.zeros(E + 8)
        k = lib.orc_construct_events(C.byref(pt), n, ob.ip(np.ascontiguousarray(g["node_pop"][l])),
                                     ob.dp(np.ascontiguousarray(g["age"][l])), m1 - m0,
                                     ob.ip(np.ascontiguousarray(np.resize(g["mig_band"][m0:m1], max(m1 - m0, 1)).astype(np.int32))),
                                     ob.ip(np.ascontiguousarray(np.resize(g["mig_target"][m0:m1], max(m1 - m0, 1)).astype(np.int32))),
                                     ob.ip(np.ascontiguousarray(np.resize(g["mig_source"][m0:m1], max(m1 - m0, 1)).astype(np.int32))),
                                     ob.dp(np.ascontiguousarray(np.resize(g["mig_age"][m0:m1], max(m1 - m0, 1)).astype(np.float64))),
                                     ob.ip(ps), ob.ip(ty), ob.ip(idd), ob.dp(el))
        assert k == E
        assert np.array_equal(ps, g["pop_start"][l])
        ref_t, ref_e = g["ev_type"][e0:e1], g["ev_time"][e0:e1]
        # zero-length intervals may order tied events differently after incremental edits: compare the
        # multiset of types per population and the cumulative times of non-tied events
        for p in range(Q):
            a, b = ps[p], ps[p + 1]
            assert sorted(ty[a:b]) == sorted(ref_t[a:b])
            assert np.allclose(np.cumsum(el[a:b])[-1], np.cumsum(ref_e[a:b])[-1], rtol=1e-12)
            assert np.allclose(np.sort(np.cumsum(el[a:b])), np.sort(np.cumsum(ref_e[a:b])), rtol=1e-9, atol=1e-12)


def test_ops_traces_match_reference_fixture():
    """Save/revert protocol (adjustGenNodeAge, executeGenSPR, scaleAllNodeAges, revertToSaved,
    resetSaved): oracle replays the recorded reference traces bit-exactly."""
    z = np.load(golden_io.os.path.join(golden_io.HERE, "golden", "ops_reference.npz"))
    for i in range(int(z["num_cases"])):
        c = {k[len(f"c{i}_"):]: z[k] for k in z.files if k.startswith(f"c{i}_")}
        n = int(c["n"])
        lc = ob.OracleLocus(n, c["chars"], c["num_phases"], c["counts"], float(c["rate"]))
        lc.set_tree(c["father"], c["left"], c["right"], c["age"], int(c["root"]))
        tr = ops.run_ops(lc, n, int(c["seed"]), int(c["steps"]), allow_leaf_age=bool(c["leaf"]), rate_moves=bool(c["ratem"]))
        lnls, trees = [], []
        for e in tr:
            if e[0] in ("init", "final-full", "rate"):
                lnls.append([e[1], np.nan])
            else:
                lnls.append([e[1], e[3]])
                trees.append(np.concatenate([e[4], e[5], e[6], [e[8]]]).astype(np.float64).tolist() + e[7].tolist())
        assert np.array_equal(np.array(lnls), c["lnls"], equal_nan=True)
        assert np.array_equal(np.array(trees), c["trees"])


def test_edge_cases():
    lib = ob.oracle()
    # JC69 edge probability: zero / negative / tiny lengths give exactly 0 (LocusDataLikelihood.c:1843-1845)
    assert lib.orc_edge_prob(0.0) == 0.0
    assert lib.orc_edge_prob(-1.0) == 0.0
    assert lib.orc_edge_prob(1e-101) == 0.0
    assert lib.orc_edge_prob(1e-3) == (1 - np.exp(-4 * 1e-3 / 3.0)) / 4.0
    # no live patterns -> 0.0 (.c:431)
    lc = ob.OracleLocus(3, np.zeros((0, 3), np.uint8), np.zeros(0, np.int32), np.zeros(0, np.int32))
    lc.set_tree([3, 3, 4, 4, -1], [-1, -1, -1, 0, 3], [-1, -1, -1, 1, 2], [0, 0, 0, 1e-3, 2e-3], 4)
    assert lc.compute(0) == 0.0
    # all-missing column: likelihood 1 => lnL 0
    lc = ob.OracleLocus(3, np.frombuffer(b"NNN", np.uint8).reshape(1, 3), np.array([1], np.int32), np.array([7], np.int32))
    lc.set_tree([3, 3, 4, 4, -1], [-1, -1, -1, 0, 3], [-1, -1, -1, 1, 2], [0, 0, 0, 1e-3, 2e-3], 4)
    assert lc.compute(0) == 0.0
    # unchanged tree with useOld=1 returns the stored value (.c:464)
    lc = ob.OracleLocus(3, np.frombuffer(b"TCT", np.uint8).reshape(1, 3), np.array([1], np.int32), np.array([5], np.int32))
    lc.set_tree([3, 3, 4, 4, -1], [-1, -1, -1, 0, 3], [-1, -1, -1, 1, 2], [0, 0, 0, 1e-3, 2e-3], 4)
    v = lc.compute(0)
    lc.reset()
    assert lc.compute(1) == v
    # bad character -> error
    with pytest.raises(ValueError):
        ob.OracleLocus(3, np.frombuffer(b"TXT", np.uint8).reshape(1, 3), np.array([1], np.int32), np.array([5], np.int32))
