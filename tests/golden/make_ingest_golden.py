"""Generates tests/golden/ingest.npz from the UNMODIFIED reference (oracle/_ref/libgphocs_ref.so, refh_ingest =
readSeqFile + processHetPatterns per locus exactly as processAlignments does, GPhoCS.c:258-440).

Run in the build container, where /root/reference exists:   python tests/golden/make_ingest_golden.py

The fixture holds, per case, the text of a random sequence file (oracle/ingest.py:random_seq_file — diploid genotypes
incl. three-way codes, N runs, lower case, samples missing from loci, unknown sample names, repeated het columns), the
sample list, and what the reference hands to initializeLocusData for every locus (CSR arrays)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ingest as oi  # noqa: E402

# name -> (sample slots, loci, seed, keyword overrides of random_seq_file)
CASES = {
    "mixed": (["h1", "h2", "d1", "", "d2", "", "h3", "d3", ""], 40, 11, {}),
    "diploid_heavy": (["a", "", "b", "", "c", "", "d", "", "e", "", "f", ""], 30, 12, dict(het=0.25, three_way=0.03)),
    "haploid": ([f"s{i}" for i in range(16)], 30, 13, dict(length=(200, 400))),
    "dense": (["x", "", "y", "", "z1", "z2", "w", ""], 6, 14, dict(length=(1500, 2500), mut=0.35, het=0.3)),
}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (names, L, seed, kw) in CASES.items():
            path = os.path.join(tmp, f"{name}.txt")
            oi.random_seq_file(path, names, L, seed, **kw)
            ref = oi.reference_ingest(path, names)
            assert ref is not None and len(ref) == L
            out[f"{name}__text"] = np.frombuffer(open(path, "rb").read(), np.uint8)
            out[f"{name}__names"] = np.array(names)
            out[f"{name}__patt_start"] = np.cumsum([0] + [len(r[1]) for r in ref]).astype(np.int64)
            out[f"{name}__unph_start"] = np.cumsum([0] + [len(r[2]) for r in ref]).astype(np.int64)
            out[f"{name}__chars"] = np.concatenate([r[0].reshape(-1, len(names)) for r in ref])
            out[f"{name}__num_phases"] = np.concatenate([r[1] for r in ref])
            out[f"{name}__counts"] = np.concatenate([r[2] for r in ref])
            print(name, "loci", L, "patterns", out[f"{name}__unph_start"][-1], "phased", out[f"{name}__patt_start"][-1],
                  "max U", max(len(r[2]) for r in ref))
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), **out)


if __name__ == "__main__":
    main()
