"""GPU: alignment ingest (header group E, k_ingest) against the golden vectors dumped from the reference, the oracle on
fresh random files, the reference's error behaviour, and end to end into the data likelihood."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

gp = importlib.import_module("g-phocs_b200")
from oracle import bindings as ob  # noqa: E402
from oracle import ingest as oi  # noqa: E402
from test_oracle_ingest import BAD_FILES, CASES, golden_case  # noqa: E402


def assert_alignment_equals(a, expected):
    """expected: list of (chars, num_phases, counts) per locus"""
    assert a.L == len(expected)
    for l, (chars, ph, cnt) in enumerate(expected):
        got = a.locus(l)
        assert np.array_equal(got[0], chars), l
        assert np.array_equal(got[1], ph), l
        assert np.array_equal(got[2], cnt), l


@pytest.mark.parametrize("name", CASES)
def test_golden_patterns_bit_exact(name, tmp_path):
    """'dense' has loci with more than 256 distinct patterns: the second table size is exercised too."""
    path, names, g = golden_case(name, tmp_path)
    a = gp.Alignment.read(path, names)
    assert np.array_equal(a.patt_start, g["patt_start"]) and np.array_equal(a.unph_start, g["unph_start"])
    assert np.array_equal(a.chars, g["chars"])
    assert np.array_equal(a.num_phases, g["num_phases"])
    assert np.array_equal(a.counts, g["counts"])
    assert a.locus_name(0) == "locus1"
    a.close()


@pytest.mark.parametrize("seed,kw,names", [
    (201, {}, ["h1", "d1", "", "d2", "", "h2", "h3", "d3", "", "d4", ""]),
    (202, dict(het=0.35, three_way=0.05), ["a", "", "b", "", "c", "", "d", "", "e", "", "f", "", "g", "", "h", ""]),
    (203, dict(missing=0.4, drop_sample=0.4), ["h1", "d1", "", "h2"]),
    (204, dict(length=(1, 6)), ["h1", "h2", "d1", ""]),
    (205, dict(mut=0.01, length=(900, 1100)), [f"s{i}" for i in range(40)]),      # three key words
    (206, dict(het=0.1), [x for i in range(12) for x in (f"d{i}", "")]),          # configs[3] shape: 12 diploids
])
def test_random_files_match_the_oracle(seed, kw, names, tmp_path):
    path = str(tmp_path / "seqs.txt")
    oi.random_seq_file(path, names, 60, seed, **kw)
    a = gp.Alignment.read(path, names)
    assert_alignment_equals(a, oi.oracle_ingest(path, names))
    # the canonical (unphased) patterns are what processLocusAlignment stores
    lib = oi._lib()
    loci = oi.parse_seq_file(path, names)
    n = len(names)
    for l in (0, len(loci) - 1):
        _, rows, length = loci[l]
        buf = C.create_string_buffer(max(1, length) * n)
        cnt = np.zeros(max(1, length), np.int32)
        U = lib.orc_locus_patterns((C.c_char_p * n)(*rows), n, length, buf, ob.ip(cnt))
        u0, u1 = a.unph_start[l], a.unph_start[l + 1]
        assert u1 - u0 == U and a.canon[u0:u1].tobytes() == buf.raw[:U * n]
    t = a.timings()
    assert t["kernel_s"] > 0 and 0 < t["raw_bytes"] <= sum(length for _, _, length in loci) * sum(1 for x in names if x)
    a.close()


def test_num_loci_to_read_and_comments(tmp_path):
    names = ["h1", "d1", ""]
    path = str(tmp_path / "s.txt")
    with open(path, "w") as f:
        f.write("# a comment line\n3\n\nlocA 3 5\nh1 acgtn\nstranger TTTTT\nd1 ACRTN\n\nlocB 2 3\nd1 NNN\nh1 NNN\nlocC 1 2\nh1 AC\n")
    a = gp.Alignment.read(path, names, num_loci_to_read=2)
    assert a.L == 2 and a.locus_name(1) == "locB"
    assert_alignment_equals(a, oi.oracle_ingest(path, names, 2))
    assert a.patt_start[2] == a.patt_start[1]          # locB: columns of N only -> no patterns
    a.close()


@pytest.mark.parametrize("case", sorted(BAD_FILES))
def test_malformed_files_are_rejected(case, tmp_path, capfd):
    path = str(tmp_path / "bad.txt")
    with open(path, "w") as f:
        f.write(BAD_FILES[case])
    with pytest.raises(ValueError):
        gp.Alignment.read(path, ["h1", "d1", ""])
    assert "Error" in capfd.readouterr().err
    with pytest.raises(ValueError):
        gp.Alignment.read(str(tmp_path / "does_not_exist.txt"), ["h1"])


def test_phasing_given_patterns_with_and_without_symmetry_breaks(tmp_path):
    """processHetPatterns alone (pattern mode of the kernel), breakSymmetries 1 and 0."""
    names = ["d1", "", "d2", "", "h1", "d3", ""]
    dip = oi.slot_is_diploid(names)
    path = str(tmp_path / "s.txt")
    oi.random_seq_file(path, names, 25, 9, het=0.3, three_way=0.04)
    full = gp.Alignment.read(path, names)
    lib = oi._lib()
    n = len(names)
    for brk in (1, 0):
        a = gp.Alignment.phase(full.unph_start, full.canon, full.counts, dip, break_symmetries=brk)
        for l in range(full.L):
            u0, u1 = full.unph_start[l], full.unph_start[l + 1]
            U = int(u1 - u0)
            cap = 1 << 16
            phased = C.create_string_buffer(cap * n)
            ph = np.zeros(cap, np.int32)
            cnt = np.ascontiguousarray(full.counts[u0:u1], np.int32)
            P = lib.orc_expand_phases(full.canon[u0:u1].tobytes(), ob.ip(cnt), U, n, bytes(1 if d else 0 for d in dip), brk,
                                      phased, ob.ip(ph), cap)
            got = a.locus(l)
            assert P == len(got[1]) and got[0].tobytes() == phased.raw[:P * n] and np.array_equal(got[1], ph[:P])
        if brk == 1:
            assert np.array_equal(a.chars, full.chars) and np.array_equal(a.num_phases, full.num_phases)
        else:
            assert a.P > full.P
        a.close()
    full.close()


def test_too_many_distinct_patterns_fails_loudly(tmp_path, capfd):
    rng = np.random.default_rng(3)
    names = [f"s{i}" for i in range(12)]
    S = 4000
    path = str(tmp_path / "s.txt")
    oi.write_seq_file(path, [("big", [(nm, "".join(rng.choice(list("TCAG"), S))) for nm in names], S)])
    with pytest.raises(ValueError):
        gp.Alignment.read(path, names)
    assert "distinct site patterns" in capfd.readouterr().err


def test_reference_entry_points_through_the_library(tmp_path):
    """readSeqFile / processHetPatterns / freeAlignmentData need the host program's AlignmentData; without it they
    refuse instead of guessing.  (With it: tests/test_gpu_dropin.py runs G-PhoCS linked without AlignmentProcessor.o.)"""
    lib = gp.lib()
    path = str(tmp_path / "s.txt")
    with open(path, "w") as f:
        f.write("1\nloc 1 4\nh1 ACGT\n")
    arr = (C.c_char_p * 1)(b"h1")
    assert lib.readSeqFile(path.encode(), 1, arr, 0) == -1


def test_store_from_alignment_gives_the_oracle_likelihood(tmp_path):
    synth = importlib.import_module("g-phocs_b200.synth")
    model = synth.config("dip8mig")
    seq = str(tmp_path / "seqs.txt")
    w = synth.generate(model, 50, seed=5, seqfile=seq)
    names = synth.sample_slots(model)
    a = gp.Alignment.read(seq, names)
    st = a.store()
    st.set_trees(w.father, w.left, w.right, w.age, w.root)
    lnl = st.evaluate(0)
    parsed = oi.parse_seq_file(seq, names)
    N = 2 * w.n - 1
    fa, le, ri, ag = (np.asarray(x).reshape(w.L, N) for x in (w.father, w.left, w.right, w.age))
    for l in range(0, w.L, 7):
        chars, ph, cnt = oi.oracle_locus(parsed[l][1], parsed[l][2], names)
        oc = ob.OracleLocus(w.n, chars, ph, cnt)
        oc.set_tree(fa[l], le[l], ri[l], ag[l], int(w.root[l]))
        ref = oc.compute(0)
        assert abs(lnl[l] - ref) <= 1e-10 * abs(ref), (l, lnl[l], ref)
    st.close(); a.close()
