"""Checkers for the pattern / phase producer.  TEST INFRASTRUCTURE ONLY (see bindings.py).

parse_seq_file   : the sequence-file format of readSeqFile / readSeqs (AlignmentProcessor.c:468-860), restated
oracle_ingest    : parse + oracle/ingest_oracle.c per locus
reference_ingest : the unmodified reference through oracle/ref_harness.c:refh_ingest
write_seq_file   : writer for test inputs
Each ingest returns a list of (chars uint8 [P][n], num_phases int32 [P], counts int32 [U]) per locus — the arguments
processAlignments passes to initializeLocusData (GPhoCS.c:403)."""
import ctypes as C

import numpy as np

from . import bindings as ob

LEGAL = set("TCAGYWKMSRVDBHN")
PARTIAL = set("YWKMSRVDBH")


class SeqFileError(ValueError):
    pass


def slot_is_diploid(names):
    """A nameless slot makes itself and the slot before it the two haplotypes of one diploid sample (:219-226)."""
    dip = [False] * len(names)
    for s, nm in enumerate(names):
        if not nm:
            if s == 0:
                raise SeqFileError("first sample cannot be nameless")
            dip[s - 1] = dip[s] = True
    return dip


def parse_seq_file(path, names, num_loci_to_read=0):
    """-> [(locus name, rows)], rows[slot] = upper-cased bytes of that slot's sequence or None."""
    dip = slot_is_diploid(names)
    slot_of = {}
    for s, nm in enumerate(names):
        if nm and nm not in slot_of:
            slot_of[nm] = s
    with open(path) as f:
        lines = [ln for ln in (x.split("#")[0].split() for x in f) if ln]
    if not lines:
        raise SeqFileError("unexpected end of file when reading the number of loci")
    try:
        num_loci = int(lines[0][0])
    except ValueError:
        raise SeqFileError(f"expected number of loci, got {lines[0][0]}")
    if num_loci <= 0:
        raise SeqFileError("at least one locus must be specified")
    if 0 < num_loci_to_read < num_loci:
        num_loci = num_loci_to_read
    pos, loci, seen = 1, [], set()
    for l in range(num_loci):
        if pos >= len(lines):
            raise SeqFileError(f"sequence file says {num_loci} loci but holds {l}")
        head = lines[pos]
        pos += 1
        if len(head) < 3:
            raise SeqFileError(f"short header for locus {l + 1}")
        k, length = int(head[1]), int(head[2])
        if k <= 0:
            raise SeqFileError(f"locus {l + 1} has no samples")
        rows = [None] * len(names)
        for _ in range(k):
            if pos >= len(lines):
                raise SeqFileError(f"unexpected end of file in locus {l + 1}")
            tok = lines[pos]
            pos += 1
            s = slot_of.get(tok[0])
            if s is None:
                continue
            seq = tok[1].upper() if len(tok) > 1 else ""
            if len(seq) < length:
                raise SeqFileError(f"sample {tok[0]} has {len(seq)} bases instead of {length}")
            if len(seq) > length:
                raise SeqFileError(f"sample {tok[0]} might be longer than {length}")
            for i, ch in enumerate(seq):
                if ch not in LEGAL:
                    raise SeqFileError(f"illegal base {ch} at site {i + 1} of sample {tok[0]}")
                if ch in PARTIAL and not dip[s]:
                    raise SeqFileError(f"ambiguity {ch} at site {i + 1} of haploid sample {tok[0]}")
            rows[s] = seq.encode()
            seen.add(s)
        loci.append((head[0], rows, length))
    for s, nm in enumerate(names):
        if nm and s not in seen:
            raise SeqFileError(f"sample {nm} has no sequence in the file")
    return loci


def _lib():
    lib = ob.oracle()
    if not getattr(lib, "_ingest_ready", False):
        lib.orc_locus_patterns.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_char_p, ob.c_int_p]
        lib.orc_expand_phases.argtypes = [C.c_char_p, ob.c_int_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p,
                                          ob.c_int_p, C.c_int]
        lib.orc_canonize_column.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib._ingest_ready = True
    return lib


def oracle_locus(rows, length, names):
    lib = _lib()
    n = len(names)
    dip = bytes(1 if d else 0 for d in slot_is_diploid(names))
    arr = (C.c_char_p * n)(*rows)
    patterns = C.create_string_buffer(max(1, length) * n)
    counts = np.zeros(max(1, length), np.int32)
    U = lib.orc_locus_patterns(arr, n, length, patterns, ob.ip(counts))
    if U < 0:
        raise SeqFileError("illegal symbol in a column")
    cap = 4 * max(U, 1)
    while True:
        phased = C.create_string_buffer(cap * n)
        num_phases = np.zeros(cap, np.int32)
        P = lib.orc_expand_phases(patterns, ob.ip(counts), U, n, dip, 1, phased, ob.ip(num_phases), cap)
        if P >= 0:
            break
        cap = -P
    chars = np.frombuffer(phased.raw[:P * n], np.uint8).reshape(P, n).copy()
    return chars, num_phases[:P].copy(), counts[:U].copy()


def oracle_ingest(path, names, num_loci_to_read=0):
    return [oracle_locus(rows, length, names) for _, rows, length in parse_seq_file(path, names, num_loci_to_read)]


def reference_ingest(path, names, num_loci_to_read=0):
    """None if the reference rejects the file (its message goes to stderr)."""
    lib = ob.ref()
    n = len(names)
    arr = (C.c_char_p * n)(*[nm.encode() for nm in names])
    lib.refh_ingest.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int]
    L = lib.refh_ingest(path.encode(), n, arr, num_loci_to_read)
    if L < 0:
        return None
    out = []
    nl, npat, nun = C.c_int(), C.c_int(), C.c_int()
    for i in range(lib.refh_recorded_count()):
        lib.refh_recorded_dims(i, C.byref(nl), C.byref(npat), C.byref(nun))
        chars = np.zeros((max(1, npat.value), n), np.uint8)
        ph = np.zeros(max(1, npat.value), np.int32)
        cnt = np.zeros(max(1, nun.value), np.int32)
        lib.refh_recorded_get(i, chars.ctypes.data_as(C.c_char_p), ob.ip(ph), ob.ip(cnt))
        out.append((chars[:npat.value], ph[:npat.value], cnt[:nun.value]))
    return out


def write_seq_file(path, loci):
    """loci = [(name, [(sample name, sequence str)], length)]"""
    with open(path, "w") as f:
        f.write(f"{len(loci)}\n\n")
        for name, seqs, length in loci:
            f.write(f"{name} {len(seqs)} {length}\n")
            for nm, sq in seqs:
                f.write(f"{nm}\t{sq}\n")
            f.write("\n")


def random_seq_file(path, names, num_loci, seed, length=(30, 120), het=0.08, three_way=0.01, missing=0.03,
                    drop_sample=0.1, lower=0.1, mut=0.06, stranger=0.1):
    """Random alignments exercising diploid genotypes (two- and three-way codes), N runs, lower case, samples absent
    from a locus, sample names the control file does not know, all-N columns and repeated het patterns."""
    rng = np.random.default_rng(seed)
    dip = slot_is_diploid(names)
    real = [(s, nm) for s, nm in enumerate(names) if nm]
    two = {frozenset("TC"): "Y", frozenset("TA"): "W", frozenset("TG"): "K", frozenset("CA"): "M", frozenset("CG"): "S",
           frozenset("AG"): "R"}
    loci = []
    for l in range(num_loci):
        S = int(rng.integers(length[0], length[1] + 1))
        anc = rng.choice(list("TCAG"), S)
        seqs = []
        allN = rng.random(S) < 0.03
        for s, nm in real:
            if len(seqs) > 0 and rng.random() < drop_sample:
                continue
            a = np.where(rng.random(S) < mut, rng.choice(list("TCAG"), S), anc)
            if dip[s]:
                b = np.where(rng.random(S) < het, rng.choice(list("TCAG"), S), a)
                g = np.array([x if x == y else two[frozenset((x, y))] for x, y in zip(a, b)])
                g = np.where(rng.random(S) < three_way, rng.choice(list("VDBH"), S), g)
            else:
                g = a
            g = np.where((rng.random(S) < missing) | allN, "N", g)
            sq = "".join(g)
            if rng.random() < lower:
                sq = sq.lower()
            seqs.append((nm, sq))
        if rng.random() < stranger:
            seqs.insert(int(rng.integers(0, len(seqs) + 1)), ("not_in_control_file", "".join(rng.choice(list("TCAG"), S))))
        if rng.random() < 0.3 and S > 8:      # repeat a few columns so het patterns with count > 1 occur
            cols = rng.integers(0, S, 4)
            seqs = [(nm, sq + "".join(sq[c] for c in cols)) for nm, sq in seqs]
            S += 4
        loci.append((f"locus{l + 1}", seqs, S))
    # every named sample must occur somewhere (readSeqFile's final check): the first locus carries all of them
    have = {nm for nm, _ in loci[0][1]}
    S0 = loci[0][2]
    for s, nm in real:
        if nm not in have:
            loci[0][1].append((nm, "".join(rng.choice(list("TCAG"), S0))))
    write_seq_file(path, loci)
    return loci
