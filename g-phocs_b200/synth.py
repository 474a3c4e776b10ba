"""Synthetic workloads of the shapes named in BASELINE.json `configs`.

Thin ctypes wrapper over csrc/synth.cpp (libgphocs_synth.so) plus a writer for the reference's
control-file format (/root/reference/src/MCMCcontrol.c:575-1345) so the very
same alignment can be pushed through the reference's own ingest when building golden fixtures.
Host-only input infrastructure: nothing here is on the measured path.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "csrc", "libgphocs_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = C.CDLL(path)
        _LIB.synth_create.restype = C.c_void_p
        _LIB.synth_num_leaves.argtypes = [C.c_void_p]
        _LIB.synth_free.argtypes = [C.c_void_p]
    return _LIB


@dataclass
class Model:
    """Population tree. Current pops are 0..C-1 in `cur` order, ancestral pops follow in `anc` order
    (the reference's numbering, MCMCcontrol.c:800,942)."""
    name: str
    cur: list            # [(pop name, haploid leaves)]
    anc: list            # [(pop name, child a, child b, tau)]
    bands: list = field(default_factory=list)   # [(source, target, mig rate)]
    diploid: bool = False
    theta: float = 1e-3
    sample_age: dict = field(default_factory=dict)   # pop name -> age of its samples
    rate_shape: float = 0.0     # >0: per-locus rates ~ Gamma(shape, mean 1)  (locus-mut-rate VAR)
    sites: int = 1000

    @property
    def names(self):
        return [c[0] for c in self.cur] + [a[0] for a in self.anc]

    @property
    def numCurPops(self):
        return len(self.cur)

    @property
    def numPops(self):
        return 2 * len(self.cur) - 1

    @property
    def numLeaves(self):
        return sum(c[1] for c in self.cur)

    def arrays(self):
        names = self.names
        idx = {n: i for i, n in enumerate(names)}
        Q, Cn = self.numPops, self.numCurPops
        father = np.full(Q, -1, np.int32)
        son0 = np.full(Q, -1, np.int32)
        son1 = np.full(Q, -1, np.int32)
        age = np.zeros(Q)
        for k, (nm, a, b, tau) in enumerate(self.anc):
            p = Cn + k
            son0[p], son1[p] = idx[a], idx[b]
            father[idx[a]] = p
            father[idx[b]] = p
            age[p] = tau
        sample_age = np.array([self.sample_age.get(c[0], 0.0) for c in self.cur])
        theta = np.full(Q, self.theta)
        src = np.array([idx[b[0]] for b in self.bands], np.int32)
        tgt = np.array([idx[b[1]] for b in self.bands], np.int32)
        rate = np.array([b[2] for b in self.bands], np.float64)
        spp = np.array([c[1] for c in self.cur], np.int32)
        return dict(father=father, son0=son0, son1=son1, age=age, sample_age=sample_age, theta=theta,
                    band_src=src, band_tgt=tgt, band_rate=rate, samples_per_pop=spp)


def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Workload:
    """CSR bundle describing L loci: phased patterns, genealogies, flattened event chains."""
    model: Model
    L: int
    n: int
    patt_start: np.ndarray
    unph_start: np.ndarray
    chars: np.ndarray       # uint8 [sumP, n]
    num_phases: np.ndarray  # int32 [sumP]
    counts: np.ndarray      # int32 [sumU]
    father: np.ndarray      # int32 [L, N]
    left: np.ndarray
    right: np.ndarray
    node_pop: np.ndarray
    age: np.ndarray         # f64 [L, N]
    root: np.ndarray        # int32 [L]
    rate: np.ndarray        # f64 [L]
    ev_start: np.ndarray    # int64 [L+1]
    pop_start: np.ndarray   # int32 [L, Q+1]
    ev_type: np.ndarray
    ev_id: np.ndarray
    ev_time: np.ndarray
    mig_start: np.ndarray
    mig_branch: np.ndarray
    mig_band: np.ndarray
    mig_target: np.ndarray
    mig_source: np.ndarray
    mig_age: np.ndarray
    band_start: np.ndarray
    band_end: np.ndarray
    pops: dict              # Model.arrays()


def generate(model: Model, L: int, seed: int = 1, missing_frac: float = 0.0, seqfile: str = None,
             nthreads: int = 0) -> Workload:
    lib = _lib()
    a = model.arrays()
    B = len(model.bands)
    h = lib.synth_create(
        C.c_int(model.numCurPops), C.c_int(B), C.c_int(model.sites), C.c_int(int(model.diploid)),
        _ip(a["samples_per_pop"]), _ip(a["father"]), _ip(a["son0"]), _ip(a["son1"]), _ip(a["age"]),
        _ip(a["sample_age"]), _ip(a["theta"]), _ip(a["band_src"]), _ip(a["band_tgt"]), _ip(a["band_rate"]),
        C.c_double(model.rate_shape), C.c_double(missing_frac), C.c_int(L), C.c_uint64(seed),
        C.c_int(1 if seqfile else 0), C.c_int(nthreads))
    h = C.c_void_p(h)
    try:
        n = lib.synth_num_leaves(h)
        N = 2 * n - 1
        Q = model.numPops
        tot = np.zeros(4, np.int64)
        lib.synth_totals(h, _ip(tot))
        sp, su, se, sg = (int(x) for x in tot)
        patt_start = np.zeros(L + 1, np.int64)
        unph_start = np.zeros(L + 1, np.int64)
        chars = np.zeros((sp, n), np.uint8)
        num_phases = np.zeros(sp, np.int32)
        counts = np.zeros(su, np.int32)
        father = np.zeros((L, N), np.int32)
        left = np.zeros((L, N), np.int32)
        right = np.zeros((L, N), np.int32)
        node_pop = np.zeros((L, N), np.int32)
        age = np.zeros((L, N))
        root = np.zeros(L, np.int32)
        rate = np.zeros(L)
        lib.synth_export(h, _ip(patt_start), _ip(unph_start), _ip(chars), _ip(num_phases), _ip(counts),
                         _ip(father), _ip(left), _ip(right), _ip(node_pop), _ip(age), _ip(root), _ip(rate))
        ev_start = np.zeros(L + 1, np.int64)
        pop_start = np.zeros((L, Q + 1), np.int32)
        ev_type = np.zeros(se, np.int32)
        ev_id = np.zeros(se, np.int32)
        ev_time = np.zeros(se)
        mig_start = np.zeros(L + 1, np.int64)
        mg = max(sg, 1)
        mig_branch = np.zeros(mg, np.int32)
        mig_band = np.zeros(mg, np.int32)
        mig_target = np.zeros(mg, np.int32)
        mig_source = np.zeros(mg, np.int32)
        mig_age = np.zeros(mg)
        lib.synth_export_events(h, _ip(ev_start), _ip(pop_start), _ip(ev_type), _ip(ev_id), _ip(ev_time),
                                _ip(mig_start), _ip(mig_branch), _ip(mig_band), _ip(mig_target),
                                _ip(mig_source), _ip(mig_age))
        bs = np.zeros(max(B, 1))
        be = np.zeros(max(B, 1))
        lib.synth_band_times(h, _ip(bs), _ip(be))
        if seqfile:
            names = sample_names(model)
            arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
            r = lib.synth_write_seqfile(h, seqfile.encode(), arr)
            if r != 0:
                raise RuntimeError(f"synth_write_seqfile failed ({r})")
    finally:
        lib.synth_free(h)
    return Workload(model, L, n, patt_start, unph_start, chars, num_phases, counts, father, left, right,
                    node_pop, age, root, rate, ev_start, pop_start, ev_type, ev_id, ev_time, mig_start,
                    mig_branch[:sg], mig_band[:sg], mig_target[:sg], mig_source[:sg], mig_age[:sg],
                    bs[:B], be[:B], a)


def sample_names(model: Model):
    """Sample names in leaf order; a diploid sample owns two consecutive leaves (MCMCcontrol.c:1286-1345)."""
    out = []
    for nm, k in model.cur:
        step = 2 if model.diploid else 1
        for j in range(0, k, step):
            out.append(f"{nm.lower()}{j // step + 1}")
    return out


def sample_slots(model: Model):
    """dataSetup.sampleNames: one entry per leaf, "" for the second leaf of a diploid sample."""
    out = []
    for nm in sample_names(model):
        out += [nm, ""] if model.diploid else [nm]
    return out


FINETUNES = dict(coal_time=0.01, mig_time=0.3, theta=0.04, mig_rate=0.02, tau=0.0000008, mixing=0.003)


def write_control_file(model: Model, path: str, seqfile: str, tracefile: str, iterations: int = 0,
                       seed: int = 4242, iterations_per_log: int = 10, finetunes: dict = None,
                       mig_prior=(0.002, 0.00001)):
    """Control file for the reference program, SURVEY.md Appendix C layout."""
    ft = dict(FINETUNES)
    ft.update(finetunes or {})
    lines = ["GENERAL-INFO-START",
             f"\tseq-file\t{seqfile}", f"\ttrace-file\t{tracefile}",
             f"\tlocus-mut-rate\t{'VAR 1.0' if model.rate_shape > 0 else 'CONST'}",
             f"\trandom-seed\t{seed}", f"\tmcmc-iterations\t{iterations}",
             f"\titerations-per-log\t{iterations_per_log}", "\tlogs-per-line\t10",
             "\tfind-finetunes\tFALSE", f"\tfinetune-coal-time\t{ft['coal_time']}", f"\tfinetune-mig-time\t{ft['mig_time']}",
             f"\tfinetune-theta\t{ft['theta']}", f"\tfinetune-mig-rate\t{ft['mig_rate']}", f"\tfinetune-tau\t{ft['tau']:.10f}",
             f"\tfinetune-mixing\t{ft['mixing']}"]
    if model.rate_shape > 0:
        lines.append("\tfinetune-locus-rate\t0.3")
    lines += ["\ttau-theta-print\t10000.0", "\ttau-theta-alpha\t1.0", f"\ttau-theta-beta\t{1.0 / model.theta:.1f}",
              "\tmig-rate-print\t0.001", f"\tmig-rate-alpha\t{mig_prior[0]}", f"\tmig-rate-beta\t{mig_prior[1]:.8f}",
              "GENERAL-INFO-END", "", "CURRENT-POPS-START"]
    names = iter(sample_names(model))
    for nm, k in model.cur:
        step = 2 if model.diploid else 1
        samp = " ".join(f"{next(names)} {'d' if model.diploid else 'h'}" for _ in range(0, k, step))
        lines += ["\tPOP-START", f"\t\tname\t{nm}", f"\t\tsamples\t{samp}"]
        if nm in model.sample_age:
            lines.append(f"\t\tage\t{model.sample_age[nm]} e")
        lines.append("\tPOP-END")
    lines += ["CURRENT-POPS-END", "", "ANCESTRAL-POPS-START"]
    for nm, a, b, tau in model.anc:
        lines += ["\tPOP-START", f"\t\tname\t{nm}", f"\t\tchildren\t{a}\t{b}", f"\t\ttau-initial\t{tau}",
                  f"\t\ttau-beta\t{1.0 / tau:.1f}", "\tPOP-END"]
    lines += ["ANCESTRAL-POPS-END", "", "MIG-BANDS-START"]
    for s, t, _ in model.bands:
        lines += ["\tBAND-START", f"\t\tsource\t{s}", f"\t\ttarget\t{t}", "\tBAND-END"]
    lines += ["MIG-BANDS-END", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


# ----------------------------------------------------------------------------- BASELINE.json configs
def config(name: str) -> Model:
    """The five shapes of BASELINE.json `configs` (sizes per SURVEY.md §8 header)."""
    if name == "sample":      # configs[0]: 4 diploids, 7 pops, 1 band (sample-control-file.ctl shape)
        return Model("sample", [("A", 2), ("B", 2), ("C", 2), ("D", 2)],
                     [("AB", "A", "B", 5e-4), ("ABC", "AB", "C", 1e-3), ("root", "ABC", "D", 2e-3)],
                     bands=[("D", "B", 150.0)], diploid=True)
    if name == "hap16":       # configs[1]: 16 phased haplotypes, 4 pops, no migration
        return Model("hap16", [("A", 4), ("B", 4), ("C", 4), ("D", 4)],
                     [("AB", "A", "B", 5e-4), ("CD", "C", "D", 8e-4), ("root", "AB", "CD", 2e-3)])
    if name == "dip8mig":     # configs[2]: 8 unphased diploids, 4 pops + 2 bands
        return Model("dip8mig", [("A", 4), ("B", 4), ("C", 4), ("D", 4)],
                     [("AB", "A", "B", 5e-4), ("CD", "C", "D", 8e-4), ("root", "AB", "CD", 2e-3)],
                     bands=[("A", "B", 200.0), ("C", "AB", 150.0)], diploid=True)
    if name == "pop6mig4":    # configs[3]: 6 pops (2 diploids each -> 24 leaves) + 4 bands
        return Model("pop6mig4", [(c, 4) for c in "ABCDEF"],
                     [("AB", "A", "B", 5e-4), ("CD", "C", "D", 7e-4), ("EF", "E", "F", 9e-4),
                      ("ABCD", "AB", "CD", 1.6e-3), ("root", "ABCD", "EF", 3e-3)],
                     bands=[("A", "B", 200.0), ("C", "D", 200.0), ("E", "AB", 120.0), ("CD", "EF", 100.0)],
                     diploid=True)
    if name == "ancient":     # configs[4]: ancient samples + per-locus rate variation
        return Model("ancient", [("A", 4), ("B", 4), ("C", 4), ("D", 4)],
                     [("AB", "A", "B", 5e-4), ("CD", "C", "D", 8e-4), ("root", "AB", "CD", 2e-3)],
                     sample_age={"B": 2e-4}, rate_shape=4.0)
    if name == "dense":       # dense-pattern stress variant: theta = 2e-2
        m = config("hap16")
        m.name = "dense"
        m.theta = 2e-2
        m.anc = [(a, b, c, t * 20) for a, b, c, t in m.anc]
        return m
    raise KeyError(name)


CONFIG_LOCI = {"sample": 1000, "hap16": 10_000, "dip8mig": 10_000, "pop6mig4": 100_000, "ancient": 50_000}
