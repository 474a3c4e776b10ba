"""CPU: the formulation the migration kernels use (g-phocs_b200/csrc/sampler_mig.cuh) against the oracle's event-chain
statistics.

The reference keeps, per population, a time-ordered chain of events and integrates n(n-1) and n over it
(computeGenetreeStats / recalcStats, patch.c:2330-2513).  The kernels never order events: every branch is cut into
SEGMENTS (population, from, to) at population ends and at its migration events, and
    coal_stats[p] = 2 * sum over unordered pairs of segments in p of their overlap        ( = integral of n(n-1) )
    mig_stats[b]  = sum over segments in target(b) of the overlap with the band's live interval   ( = integral of n )
This file restates the two device functions that do it —
  * smgBuildSegments: 32 lanes walk 32 branches at once, one segment per lane and round, the lanes that have one take
    consecutive slots by ballot (every branch is walked once, no prefix pass);
  * smgPopStats: the segments of one population compacted into an index list and paired around a circle, entry a with
    the (count-1)/2 entries after it (and the one opposite, for an even count and a in the first half), so that every
    lane has the same number of partners —
in Python and demands: the segment list is a permutation of what a plain per-branch walk gives; the circular pairing
visits every unordered pair exactly once; statistics equal the oracle's to checkGtreeStructure's 1e-10
(patch.c:2986) on genealogies with migration events of three shapes."""
import importlib
import itertools

import numpy as np
import pytest

synth = importlib.import_module("g-phocs_b200.synth")
from oracle import bindings as ob  # noqa: E402

INF = 1e300


def walk_branch(x, father, age, node_pop, pop_father, tau, migs, band_src, band_tgt):
    """the plain walk: branch x from its node up to its father (for ever above the root), up at population ends,
    sideways (target -> source) at its migration events"""
    pop, t = int(node_pop[x]), float(age[x])
    t_end = float(age[father[x]]) if father[x] >= 0 else INF
    mine = sorted((a, b) for br, b, a in migs if br == x)
    out = []
    for _ in range(200):
        m_age, m_band = mine[0] if mine else (INF, -1)
        pop_end = float(tau[pop_father[pop]]) if pop_father[pop] >= 0 else INF
        t_next = min(t_end, pop_end, m_age)
        out.append((pop, t, t_next, x))
        if mine and m_age <= t_end and m_age <= pop_end:
            assert pop == band_tgt[m_band]
            pop, t = int(band_src[m_band]), m_age
            mine.pop(0)
            continue
        if t_end <= pop_end or pop_father[pop] < 0:
            break
        pop, t = int(pop_father[pop]), t_next
    return out


def build_segments_by_ballot(N, father, age, node_pop, pop_father, tau, migs, band_src, band_tgt):
    """smgBuildSegments: rounds of 32 branches; in every step of a round the live lanes report their next segment and
    take consecutive slots in lane order"""
    segs = []
    for x0 in range(0, N, 32):
        state = {}
        for lane in range(32):
            x = x0 + lane
            if x < N:
                state[lane] = dict(pop=int(node_pop[x]), t=float(age[x]), x=x,
                                   t_end=float(age[father[x]]) if father[x] >= 0 else INF,
                                   mine=sorted((a, b) for br, b, a in migs if br == x))
        while state:
            for lane in sorted(state):   # ballot order = lane order
                s = state[lane]
                m_age, m_band = s["mine"][0] if s["mine"] else (INF, -1)
                pop_end = float(tau[pop_father[s["pop"]]]) if pop_father[s["pop"]] >= 0 else INF
                t_next = min(s["t_end"], pop_end, m_age)
                segs.append((s["pop"], s["t"], t_next, s["x"]))
                if s["mine"] and m_age <= s["t_end"] and m_age <= pop_end:
                    s["pop"], s["t"] = int(band_src[m_band]), m_age
                    s["mine"].pop(0)
                elif s["t_end"] <= pop_end or pop_father[s["pop"]] < 0:
                    del state[lane]
                else:
                    s["pop"], s["t"] = int(pop_father[s["pop"]]), t_next
    return segs


def circular_pairs(count):
    """smgPopStats: the partners of list entry a"""
    half = (count - 1) >> 1
    for a in range(count):
        partners = half + (1 if count % 2 == 0 and a < count // 2 else 0)
        b = a
        for _ in range(partners):
            b = 0 if b + 1 == count else b + 1
            yield a, b


@pytest.mark.parametrize("count", list(range(0, 70)))
def test_circular_pairing_visits_every_pair_once(count):
    got = sorted(tuple(sorted(p)) for p in circular_pairs(count))
    assert got == sorted(itertools.combinations(range(count), 2))
    per_entry = np.bincount([a for a, _ in circular_pairs(count)], minlength=max(count, 1))[:count]
    assert count < 2 or per_entry.max() - per_entry.min() <= 1      # balanced: the point of the circle


@pytest.mark.parametrize("cfg,L", [("dip8mig", 120), ("pop6mig4", 60), ("sample", 150)])
def test_segment_formulation_reproduces_the_event_chain_statistics(cfg, L):
    w = synth.generate(synth.config(cfg), L, seed=12)
    assert len(w.mig_age) > 0
    pops = w.pops
    Q, C, B = len(pops["father"]), len(pops["samples_per_pop"]), len(pops["band_src"])
    tau = np.array(pops["age"], float)
    tau[:C] = pops["sample_age"]
    birth = np.where(np.arange(Q) < C, 0.0, tau)          # a current population exists from time 0 (PopulationTree.c:448)
    end = np.array([tau[f] if f >= 0 else INF for f in pops["father"]])
    pt, keep = ob.make_poptree(pops, w.band_start, w.band_end)
    N = w.father.shape[1]
    with_migs = 0
    for l in range(L):
        m0, m1 = int(w.mig_start[l]), int(w.mig_start[l + 1])
        migs = [(int(w.mig_branch[k]), int(w.mig_band[k]), float(w.mig_age[k])) for k in range(m0, m1)]
        with_migs += bool(migs)
        args = (w.father[l], w.age[l], w.node_pop[l], pops["father"], tau, migs, pops["band_src"], pops["band_tgt"])
        segs = build_segments_by_ballot(N, *args)
        plain = [s for x in range(N) for s in walk_branch(x, *args)]
        assert sorted(segs) == sorted(plain)
        coal, mig = np.zeros(Q), np.zeros(B)
        for p in range(Q):
            lst = [i for i, s in enumerate(segs) if s[0] == p]
            c = 0.0
            for a, b in circular_pairs(len(lst)):
                sa, sb = segs[lst[a]], segs[lst[b]]
                c += max(0.0, min(sa[2], sb[2]) - max(sa[1], sb[1]))
            coal[p] = 2.0 * c
            for b in range(B):
                if pops["band_tgt"][b] != p:
                    continue
                src = int(pops["band_src"][b])
                s0, s1 = max(birth[src], birth[p]), min(end[src], end[p])
                mig[b] = sum(max(0.0, min(s1, segs[i][2]) - max(s0, segs[i][1])) for i in lst)
        e0, e1 = int(w.ev_start[l]), int(w.ev_start[l + 1])
        _, cs, nc, ms, nm, _ = ob.oracle_gen_locus(pt, w.pop_start[l], w.ev_type[e0:e1], w.ev_id[e0:e1], w.ev_time[e0:e1])
        assert np.allclose(coal, cs, rtol=1e-10, atol=1e-15), (l, coal, cs)
        assert np.allclose(mig, ms[:B], rtol=1e-10, atol=1e-15), (l, mig, ms)
        assert np.array_equal(np.bincount(w.node_pop[l][w.n:], minlength=Q)[:Q], nc)
        assert np.array_equal(np.bincount([b for _, b, _ in migs], minlength=B)[:B], nm[:B])
    assert with_migs > L // 10
