"""CPU: the ingest restatement (oracle/ingest_oracle.c + oracle/ingest.py) against the golden vectors dumped from the
reference (tests/golden/ingest.npz) and, where the compiled reference is present, against the reference itself on
fresh random sequence files and on malformed ones."""
import os

import numpy as np
import pytest

from oracle import bindings as ob
from oracle import ingest as oi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = np.load(os.path.join(HERE, "golden", "ingest.npz"))
CASES = sorted({k.split("__")[0] for k in GOLDEN.files})


def golden_case(name, tmp_path):
    path = str(tmp_path / f"{name}.txt")
    with open(path, "wb") as f:
        f.write(GOLDEN[f"{name}__text"].tobytes())
    names = [str(x) for x in GOLDEN[f"{name}__names"]]
    g = {k: GOLDEN[f"{name}__{k}"] for k in ("patt_start", "unph_start", "chars", "num_phases", "counts")}
    return path, names, g


def assert_same(got, g):
    """got: list of (chars, num_phases, counts) per locus; g: CSR golden arrays"""
    assert len(got) == len(g["patt_start"]) - 1
    for l, (chars, ph, cnt) in enumerate(got):
        a, b, c, d = g["patt_start"][l], g["patt_start"][l + 1], g["unph_start"][l], g["unph_start"][l + 1]
        assert np.array_equal(chars, g["chars"][a:b]), l
        assert np.array_equal(ph, g["num_phases"][a:b]), l
        assert np.array_equal(cnt, g["counts"][c:d]), l


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_golden_patterns(name, tmp_path):
    path, names, g = golden_case(name, tmp_path)
    assert_same(oi.oracle_ingest(path, names), g)


def test_canonical_form_is_the_jc_orbit_minimum():
    """Independent check of cannonizeJCpattern's contract: relabelling the bases of a column never changes its canonical
    form, and the form is a fixed point."""
    import itertools
    lib = oi._lib()
    rng = np.random.default_rng(5)
    alphabet = "TCAGYWKMSRVDBHN"
    pairs = {"Y": "TC", "W": "TA", "K": "TG", "M": "CA", "S": "CG", "R": "AG"}
    triple = {"V": "T", "D": "C", "B": "A", "H": "G"}      # the excluded base
    def relabel(col, perm):
        m = dict(zip("TCAG", perm))
        out = []
        for ch in col:
            if ch in m:
                out.append(m[ch])
            elif ch in pairs:
                want = {m[pairs[ch][0]], m[pairs[ch][1]]}
                out.append(next(k for k, v in pairs.items() if set(v) == want))
            elif ch in triple:
                out.append(next(k for k, v in triple.items() if v == m[triple[ch]]))
            else:
                out.append(ch)
        return "".join(out)
    def canon(col):
        out = bytes(len(col))
        buf = (oi.C.c_char * len(col)).from_buffer_copy(out)
        assert lib.orc_canonize_column(col.encode(), buf, len(col)) == 0
        return buf.raw.decode()
    for _ in range(60):
        col = "".join(rng.choice(list(alphabet), 9))
        base = canon(col)
        assert canon(base) == base
        for perm in itertools.permutations("TCAG"):
            assert canon(relabel(col, perm)) == base, (col, perm)


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgphocs_ref.so not built")
@pytest.mark.parametrize("seed,kw", [(101, {}), (102, dict(het=0.3, three_way=0.05)), (103, dict(missing=0.4, drop_sample=0.4)),
                                     (104, dict(length=(1, 6))), (105, dict(mut=0.4, het=0.4, length=(300, 600)))])
def test_oracle_equals_the_reference_on_random_files(seed, kw, tmp_path):
    names = ["h1", "d1", "", "d2", "", "h2", "h3", "d3", "", "d4", ""]
    path = str(tmp_path / "seqs.txt")
    oi.random_seq_file(path, names, 30, seed, **kw)
    ref = oi.reference_ingest(path, names)
    got = oi.oracle_ingest(path, names)
    assert ref is not None and len(ref) == len(got)
    for l, (x, y) in enumerate(zip(got, ref)):
        for k in range(3):
            assert np.array_equal(x[k], y[k]), (l, k)


BAD_FILES = {
    "illegal_base": "1\nloc 2 4\nh1 ACGT\nd1 ACXT\n",
    "ambiguity_in_haploid": "1\nloc 2 4\nh1 ACRT\nd1 ACGT\n",
    "short_sequence": "1\nloc 2 4\nh1 ACG\nd1 ACGT\n",
    "long_sequence": "1\nloc 2 4\nh1 ACGTA\nd1 ACGT\n",
    "missing_locus": "2\nloc 2 4\nh1 ACGT\nd1 ACGT\n",
    "sample_never_seen": "1\nloc 1 4\nh1 ACGT\n",
    "no_loci": "0\n",
}


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgphocs_ref.so not built")
@pytest.mark.parametrize("case", sorted(BAD_FILES))
def test_malformed_files_are_rejected_like_the_reference(case, tmp_path):
    names = ["h1", "d1", ""]
    path = str(tmp_path / "bad.txt")
    with open(path, "w") as f:
        f.write(BAD_FILES[case])
    assert oi.reference_ingest(path, names) is None
    with pytest.raises(oi.SeqFileError):
        oi.parse_seq_file(path, names)
