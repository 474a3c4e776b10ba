"""ctypes bindings for the checkers: oracle/liboracle.so (our CPU restatement) and
oracle/_ref/libgphocs_ref.so (the unmodified reference + ref_harness.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never by the product package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgphocs_ref.so")
REF_BIN = os.path.join(_HERE, "_ref", "G-PhoCS-ref")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_int_p)


def dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dbl_p)


class OrcPopTree(C.Structure):
    _fields_ = [("numPops", C.c_int), ("numCurPops", C.c_int), ("numBands", C.c_int), ("rootPop", C.c_int),
                ("theta", c_dbl_p), ("age", c_dbl_p), ("sampleAge", c_dbl_p), ("father", c_int_p),
                ("son0", c_int_p), ("son1", c_int_p), ("samplesPerPop", c_int_p), ("bandSource", c_int_p),
                ("bandTarget", c_int_p), ("bandRate", c_dbl_p), ("bandStart", c_dbl_p), ("bandEnd", c_dbl_p)]


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError(f"{ORACLE_SO} missing: run __graft_entry__.build()")
        lib = C.CDLL(ORACLE_SO)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_int]
        lib.orc_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, c_int_p, c_int_p]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_set_rate.argtypes = [C.c_void_p, C.c_double]
        lib.orc_get_rate.argtypes = [C.c_void_p]
        lib.orc_get_rate.restype = C.c_double
        lib.orc_set_tree.argtypes = [C.c_void_p, c_int_p, c_int_p, c_int_p, c_dbl_p, C.c_int]
        lib.orc_get_tree.argtypes = [C.c_void_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p]
        lib.orc_compute.argtypes = [C.c_void_p, C.c_int]
        lib.orc_compute.restype = C.c_double
        lib.orc_get_lnl.argtypes = [C.c_void_p]
        lib.orc_get_lnl.restype = C.c_double
        lib.orc_adjust_age.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib.orc_scale_all.argtypes = [C.c_void_p, C.c_double]
        lib.orc_scale_all.restype = C.c_double
        lib.orc_spr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        lib.orc_revert.argtypes = [C.c_void_p]
        lib.orc_reset.argtypes = [C.c_void_p]
        lib.orc_check.argtypes = [C.c_void_p]
        lib.orc_get_clv.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dbl_p]
        lib.orc_edge_prob.argtypes = [C.c_double]
        lib.orc_edge_prob.restype = C.c_double
        lib.orc_construct_events.argtypes = [C.POINTER(OrcPopTree), C.c_int, c_int_p, c_dbl_p, C.c_int, c_int_p,
                                             c_int_p, c_int_p, c_dbl_p, c_int_p, c_int_p, c_int_p, c_dbl_p]
        lib.orc_gen_stats.argtypes = [C.POINTER(OrcPopTree), c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p,
                                      c_int_p, c_dbl_p, c_int_p]
        lib.orc_gen_lnl.argtypes = [C.POINTER(OrcPopTree), c_dbl_p, c_int_p, c_dbl_p, c_int_p]
        lib.orc_gen_lnl.restype = C.c_double
        _oracle = lib
    return _oracle


class OracleLocus:
    """One locus driven through the oracle; mirrors the reference's LocusData call surface."""

    def __init__(self, n, chars, num_phases, counts, rate=1.0):
        self.lib = oracle()
        self.n, self.N, self.P = n, 2 * n - 1, len(num_phases)
        self.h = C.c_void_p(self.lib.orc_create(n))
        chars = np.ascontiguousarray(chars, np.uint8)
        r = self.lib.orc_init(self.h, chars.tobytes(), self.P, ip(np.ascontiguousarray(num_phases, np.int32)),
                              ip(np.ascontiguousarray(counts, np.int32)))
        if r != 0:
            raise ValueError("orc_init failed")
        self.lib.orc_set_rate(self.h, rate)

    def set_tree(self, father, left, right, age, root):
        f, l, r = (np.ascontiguousarray(x, np.int32) for x in (father, left, right))
        self.lib.orc_set_tree(self.h, ip(f), ip(l), ip(r), dp(np.ascontiguousarray(age, np.float64)), int(root))

    def tree(self):
        f, l, r = (np.zeros(self.N, np.int32) for _ in range(3))
        a = np.zeros(self.N)
        root = C.c_int()
        self.lib.orc_get_tree(self.h, ip(f), ip(l), ip(r), dp(a), C.byref(root))
        return f, l, r, a, root.value

    def compute(self, use_old):
        return self.lib.orc_compute(self.h, int(use_old))

    def lnl(self):
        return self.lib.orc_get_lnl(self.h)

    def set_rate(self, r):
        self.lib.orc_set_rate(self.h, r)

    def adjust_age(self, node, age):
        return self.lib.orc_adjust_age(self.h, node, age)

    def scale_all(self, f):
        return self.lib.orc_scale_all(self.h, f)

    def spr(self, sub, target, age):
        return self.lib.orc_spr(self.h, sub, target, age)

    def revert(self):
        return self.lib.orc_revert(self.h)

    def reset(self):
        return self.lib.orc_reset(self.h)

    def check(self):
        return self.lib.orc_check(self.h)

    def clv(self, node, saved=False):
        out = np.zeros(4 * self.P)
        self.lib.orc_get_clv(self.h, node, int(saved), dp(out))
        return out.reshape(self.P, 4)

    def __del__(self):
        try:
            self.lib.orc_free(self.h)
        except Exception:
            pass


def make_poptree(pops, band_start, band_end):
    """OrcPopTree from Model.arrays() (synth.py); returns (struct, keepalive list)."""
    Q = len(pops["father"])
    Cn = len(pops["samples_per_pop"])
    B = len(pops["band_src"])
    keep = dict(theta=np.ascontiguousarray(pops["theta"], np.float64), age=np.ascontiguousarray(pops["age"], np.float64),
                sample_age=np.ascontiguousarray(pops["sample_age"], np.float64),
                father=np.ascontiguousarray(pops["father"], np.int32), son0=np.ascontiguousarray(pops["son0"], np.int32),
                son1=np.ascontiguousarray(pops["son1"], np.int32),
                spp=np.ascontiguousarray(pops["samples_per_pop"], np.int32),
                bsrc=np.ascontiguousarray(np.resize(pops["band_src"], max(B, 1)) if B else np.zeros(1), np.int32),
                btgt=np.ascontiguousarray(np.resize(pops["band_tgt"], max(B, 1)) if B else np.zeros(1), np.int32),
                brate=np.ascontiguousarray(np.resize(pops["band_rate"], max(B, 1)) if B else np.zeros(1), np.float64),
                bstart=np.ascontiguousarray(np.resize(band_start, max(B, 1)) if B else np.zeros(1), np.float64),
                bend=np.ascontiguousarray(np.resize(band_end, max(B, 1)) if B else np.zeros(1), np.float64))
    root = int(np.where(keep["father"] < 0)[0][0])
    pt = OrcPopTree(Q, Cn, B, root, dp(keep["theta"]), dp(keep["age"]), dp(keep["sample_age"]), ip(keep["father"]),
                    ip(keep["son0"]), ip(keep["son1"]), ip(keep["spp"]), ip(keep["bsrc"]), ip(keep["btgt"]),
                    dp(keep["brate"]), dp(keep["bstart"]), dp(keep["bend"]))
    return pt, keep


def oracle_gen_locus(pt, pop_start, ev_type, ev_id, ev_time):
    """(num_lineages, coal_stats, num_coals, mig_stats, num_migs, lnL) of one flattened genealogy."""
    lib = oracle()
    Q, B = pt.numPops, pt.numBands
    E = len(ev_type)
    nl = np.zeros(max(E, 1), np.int32)
    cs, nc = np.zeros(Q), np.zeros(Q, np.int32)
    ms, nm = np.zeros(max(B, 1)), np.zeros(max(B, 1), np.int32)
    rc = lib.orc_gen_stats(C.byref(pt), ip(np.ascontiguousarray(pop_start, np.int32)),
                           ip(np.ascontiguousarray(ev_type, np.int32)), ip(np.ascontiguousarray(ev_id, np.int32)),
                           dp(np.ascontiguousarray(ev_time, np.float64)), ip(nl), dp(cs), ip(nc), dp(ms), ip(nm))
    if rc != 0:
        raise ValueError("malformed event chain")
    lnl = lib.orc_gen_lnl(C.byref(pt), dp(cs), ip(nc), dp(ms), ip(nm))
    return nl[:E], cs, nc, ms[:B], nm[:B], lnl


# --------------------------------------------------------------------------------------- reference
_ref = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """The compiled reference + harness. Global state: refh_setup may be called once per process."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError(f"{REF_SO} missing (built only where /root/reference exists)")
        lib = C.CDLL(REF_SO)
        for name in ("refh_data_lnl", "refh_compute_data_lnl", "refh_total_data_lnl", "refh_stored_gen_lnl",
                     "refh_recompute_gen", "refh_gen_lnl", "refh_time_data_full", "refh_time_gen_full",
                     "refh_time_both_once", "computeLocusDataLikelihood", "getLocusDataLikelihood",
                     "scaleAllNodeAges", "getNodeAge", "getLocusMutationRate"):
            getattr(lib, name).restype = C.c_double
        lib.refh_locus.restype = C.c_void_p
        lib.createLocusData.restype = C.c_void_p
        lib.createLocusData.argtypes = [C.c_int, C.c_ushort]
        lib.freeLocusData.argtypes = [C.c_void_p]
        lib.computeLocusDataLikelihood.argtypes = [C.c_void_p, C.c_ushort]
        lib.getLocusDataLikelihood.argtypes = [C.c_void_p]
        lib.adjustGenNodeAge.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib.scaleAllNodeAges.argtypes = [C.c_void_p, C.c_double]
        lib.executeGenSPR.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        lib.revertToSaved.argtypes = [C.c_void_p]
        lib.resetSaved.argtypes = [C.c_void_p]
        lib.checkLocusDataLikelihood.argtypes = [C.c_void_p]
        lib.setLocusMutationRate.argtypes = [C.c_void_p, C.c_double]
        lib.getLocusMutationRate.argtypes = [C.c_void_p]
        lib.getLocusRoot.argtypes = [C.c_void_p]
        lib.refh_set_tree.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, c_dbl_p, C.c_int]
        lib.refh_init_locus.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, c_int_p, c_int_p]
        lib.refh_get_locus_tree.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p]
        lib.refh_time_data_full.argtypes = [C.c_int, c_dbl_p]
        lib.refh_time_gen_full.argtypes = [C.c_int, c_dbl_p]
        lib.refh_time_both_once.argtypes = [c_dbl_p, c_dbl_p]
        lib.refh_set_theta.argtypes = [C.c_int, C.c_double]
        lib.refh_set_mig_rate.argtypes = [C.c_int, C.c_double]
        _ref = lib
    return _ref


class RefLocus:
    """One stand-alone locus driven through the reference's own LocusData API."""

    def __init__(self, n, chars, num_phases, counts, rate=1.0):
        self.lib = ref()
        self.n, self.N, self.P = n, 2 * n - 1, len(num_phases)
        self.h = C.c_void_p(self.lib.createLocusData(n, 1))
        chars = np.ascontiguousarray(chars, np.uint8)
        r = self.lib.refh_init_locus(self.h, n, chars.tobytes(), self.P,
                                     ip(np.ascontiguousarray(num_phases, np.int32)),
                                     ip(np.ascontiguousarray(counts, np.int32)))
        if r != 0:
            raise ValueError("initializeLocusData failed")
        self.lib.setLocusMutationRate(self.h, rate)

    def set_tree(self, father, left, right, age, root):
        f, l, r = (np.ascontiguousarray(x, np.int32).copy() for x in (father, left, right))
        a = np.ascontiguousarray(age, np.float64).copy()
        self.lib.refh_set_tree(self.h, self.n, ip(f), ip(l), ip(r), dp(a), int(root))

    def tree(self):
        f, l, r = (np.zeros(self.N, np.int32) for _ in range(3))
        a = np.zeros(self.N)
        root = C.c_int()
        self.lib.refh_get_locus_tree(self.h, self.n, ip(f), ip(l), ip(r), dp(a), C.byref(root))
        return f, l, r, a, root.value

    def compute(self, use_old):
        return self.lib.computeLocusDataLikelihood(self.h, int(use_old))

    def lnl(self):
        return self.lib.getLocusDataLikelihood(self.h)

    def set_rate(self, r):
        self.lib.setLocusMutationRate(self.h, r)

    def adjust_age(self, node, age):
        return self.lib.adjustGenNodeAge(self.h, node, age)

    def scale_all(self, f):
        return self.lib.scaleAllNodeAges(self.h, f)

    def spr(self, sub, target, age):
        return self.lib.executeGenSPR(self.h, sub, target, age)

    def revert(self):
        return self.lib.revertToSaved(self.h)

    def reset(self):
        return self.lib.resetSaved(self.h)

    def check(self):
        return self.lib.checkLocusDataLikelihood(self.h)
