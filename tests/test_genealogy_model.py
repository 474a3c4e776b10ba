"""CPU: the formulation k_gen_eval uses (g-phocs_b200/csrc/gen_kernels.cuh) against the reference's own numbers in
the golden fixtures.

The reference walks the populations in post-order and carries the number of lineages from chain to chain
(computeGenetreeStats / recalcStats, patch.c:2330-2513).  The kernel walks all chains of a locus concurrently, so it
first needs the lineages each chain starts with: pass A sums the lineage change of every chain from the event types
alone (the 4-bit table `lineageStep`), a post-order sweep adds the sons' ends, pass B then accumulates n(n-1)t and n*t
per live band chain by chain.  This file restates that in Python and demands, on every fixture dumped from the
reference: the lineages at every event `==` the reference's num_lineages, exact coalescence / migration counts, and
statistics within checkGtreeStructure's 1e-10 (patch.c:2986; the reference maintains them incrementally)."""
import numpy as np
import pytest

import golden_io

COAL, IN_MIG, OUT_MIG, BAND_START, BAND_END, SAMPLES_START, END_CHAIN, DUMMY = range(8)


def lineage_step(t):
    """lineageStep of gen_kernels.cuh: 4-bit two's complement entries of 0x000001ff, SAMPLES_START handled apart."""
    x = (0x000001FF << (28 - 4 * t)) & 0xFFFFFFFF
    x = x - (1 << 32) if x & 0x80000000 else x
    return x >> 28


def post_order(father, son0, son1):
    Q = len(father)
    root = int(np.flatnonzero(np.asarray(father) < 0)[0])
    out = []

    def rec(p):
        if son0[p] >= 0:
            rec(int(son0[p])); rec(int(son1[p]))
        out.append(p)
    rec(root)
    assert len(out) == Q
    return out


def test_lineage_step_table():
    assert [lineage_step(t) for t in range(8)] == [-1, -1, 1, 0, 0, 0, 0, 0]


@pytest.mark.parametrize("name", golden_io.CASES)
def test_two_pass_formulation_reproduces_the_reference(name):
    g = golden_io.load(name)
    Q, C, B, L = int(g["Q"]), int(g["C"]), int(g["B"]), int(g["L"])
    order = post_order(g["pop_father"], g["pop_son0"], g["pop_son1"])
    smp = g["samples_per_pop"]
    for l in range(L):
        e0 = int(g["ev_start"][l])
        ps = g["pop_start"][l].astype(int)
        ty = g["ev_type"][e0:int(g["ev_start"][l + 1])].astype(int)
        idd = g["ev_id"][e0:int(g["ev_start"][l + 1])].astype(int)
        tm = g["ev_time"][e0:int(g["ev_start"][l + 1])]
        ref_n = g["ev_lineages"][e0:int(g["ev_start"][l + 1])].astype(int)
        # pass A: lineage change of every chain, from the event types alone
        delta = np.zeros(Q, int)
        for p in range(Q):
            for e in range(ps[p], ps[p + 1]):
                delta[p] += smp[p] if ty[e] == SAMPLES_START else lineage_step(ty[e])
        # post-order sweep: an ancestral population starts with what its sons end with (patch.c:2336-2347)
        start, end = np.zeros(Q, int), np.zeros(Q, int)
        for p in order:
            start[p] = end[g["pop_son0"][p]] + end[g["pop_son1"][p]] if p >= C else 0
            end[p] = start[p] + delta[p]
        # pass B: statistics chain by chain, in the reference's order of operations (patch.c:2403-2486)
        coal, ncoal = np.zeros(Q), np.zeros(Q, int)
        mig, nmig = np.zeros(max(B, 1)), np.zeros(max(B, 1), int)
        for p in range(Q):
            n, live = start[p], []
            for e in range(ps[p], ps[p + 1]):
                assert n == ref_n[e], (name, l, p, e)
                coal[p] += n * (n - 1) * tm[e]
                for b in live:
                    mig[b] += n * tm[e]
                t = ty[e]
                if t == IN_MIG:
                    nmig[idd[e]] += 1
                elif t == BAND_START:
                    live.append(idd[e]); mig[idd[e]] = 0.0; nmig[idd[e]] = 0
                elif t == BAND_END:
                    live.remove(idd[e])
                ncoal[p] += t == COAL
                n += smp[p] if t == SAMPLES_START else lineage_step(t)
            assert n == end[p] and not live
        assert end[order[-1]] == 1                      # one lineage leaves the root population
        assert np.array_equal(ncoal, g["num_coals"][l])
        assert np.allclose(coal, g["coal_stats"][l], rtol=1e-10, atol=1e-14)
        if B:
            assert np.array_equal(nmig[:B], g["num_migs"][l])
            assert np.allclose(mig[:B], g["mig_stats"][l], rtol=1e-10, atol=1e-14)
