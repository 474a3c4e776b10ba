"""gphocs-b200: B200-native per-locus likelihood path of G-PhoCS.

The product is csrc/libgphocs_b200.so (hand-written CUDA for sm_100a behind the C ABI of
include/gphocs_b200.h).  This package is the thin Python binding used by tests and bench.py; it holds
no arithmetic.  There is no CPU fallback: constructing a store without the built library or without a
CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPHOCS_B200_LIB") or os.path.join(_HERE, "csrc", "libgphocs_b200.so")   # the variable: development builds

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_ll_p = C.POINTER(C.c_longlong)

OP_ADJUST_AGE, OP_SPR, OP_SCALE_ALL, OP_COMMIT, OP_REVERT, OP_SET_RATE = range(6)

OP_DTYPE = np.dtype([("locus", np.int32), ("type", np.int32), ("a", np.int32), ("b", np.int32), ("x", np.float64)])


class GenericBinaryTree(C.Structure):   # layout of src/GenericTree.h:29-39
    _fields_ = [("numLeaves", C.c_int), ("rootId", C.c_int), ("leafNames", C.c_void_p), ("father", c_int_p),
                ("leftSon", c_int_p), ("rightSon", c_int_p), ("label1", c_dbl_p), ("label2", c_dbl_p)]


_lib = None


def lib():
    """Loads libgphocs_b200.so and declares every entry point of include/gphocs_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, ci, cd, cus = C.c_void_p, C.c_int, C.c_double, C.c_ushort
    sig = {
        # A. LocusData surface
        "createLocusData": (vp, [ci, cus]),
        "initializeLocusData": (ci, [vp, C.POINTER(C.c_char_p), ci, c_int_p, c_int_p]),
        "freeLocusData": (ci, [vp]),
        "attachLeaf_UNUSED": (ci, [vp, ci, ci, cd]),
        "setLocusMutationRate": (None, [vp, cd]),
        "getLocusMutationRate": (cd, [vp]),
        "computeAllConditionals": (ci, [vp]),
        "computeLocusDataLikelihood": (cd, [vp, cus]),
        "computePatternLogLikelihood": (cd, [vp, ci, c_int_p, c_int_p]),
        "computeLocusDataLikelihood_deb": (cd, [vp, cus]),
        "addSitePatterns": (cd, [vp, ci, c_int_p, c_int_p, cus]),
        "reduceSitePatterns": (cd, [vp, ci, c_int_p, c_int_p, cus]),
        "checkLocusDataLikelihood": (ci, [vp]),
        "revertToSaved": (ci, [vp]),
        "resetSaved": (ci, [vp]),
        "adjustGenNodeAge": (ci, [vp, ci, cd]),
        "scaleAllNodeAges": (cd, [vp, cd]),
        "executeGenSPR": (ci, [vp, ci, ci, cd]),
        "copyGenericTreeToLocus": (ci, [vp, C.POINTER(GenericBinaryTree)]),
        "printLocusGenTree": (None, [vp, vp, c_int_p, c_int_p]),
        "printLocusDataStats": (None, [vp, ci]),
        "printLocusDataPatterns": (None, [vp, vp]),
        "computePairwiseLCAs": (ci, [vp, C.POINTER(c_int_p), c_int_p]),
        "getSortedAges": (ci, [vp, c_dbl_p]),
        "getLocusDataLikelihood": (cd, [vp]),
        "getLocusRoot": (ci, [vp]),
        "getNodeAge": (cd, [vp, ci]),
        "getNodeFather": (ci, [vp, ci]),
        "getNodeSon": (ci, [vp, ci, cus]),
        "gpuLociStore": (vp, []),
        "gpuLocusIndex": (ci, [vp]),
        # B. batched engine
        "gphocsStoreCreate": (vp, [ci, ci, ci, c_ll_p, c_ll_p, C.c_char_p, c_int_p, c_int_p]),
        "gphocsStoreDestroy": (ci, [vp]),
        "gphocsStoreSetStream": (ci, [vp, vp]),
        "gphocsStoreNumLoci": (ci, [vp]),
        "gphocsStoreNumLeaves": (ci, [vp]),
        "gphocsStoreNumColumns": (C.c_longlong, [vp]),
        "gphocsStoreDeviceBytes": (C.c_longlong, [vp]),
        "gphocsStoreSetTrees": (ci, [vp, ci, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p]),
        "gphocsStoreSetTreesPacked": (ci, [vp, ci, C.POINTER(C.c_short), c_dbl_p, c_int_p]),
        "gphocsStoreGetTrees": (ci, [vp, ci, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p]),
        "gphocsStoreSetRates": (ci, [vp, ci, c_int_p, c_dbl_p]),
        "gphocsStoreApplyOps": (ci, [vp, ci, vp, c_int_p]),
        "gphocsStoreApplyOpsAsync": (ci, [vp, ci, vp]),
        "gphocsStoreEvaluate": (ci, [vp, ci, c_int_p, ci, c_dbl_p, c_dbl_p]),
        "gphocsStoreEvaluateDevice": (ci, [vp, ci, C.POINTER(vp), C.POINTER(vp)]),
        "gphocsStoreGetLnL": (ci, [vp, ci, c_int_p, c_dbl_p]),
        "gphocsStoreGetRates": (ci, [vp, ci, c_int_p, c_dbl_p]),
        "gphocsStoreGetClv": (ci, [vp, ci, ci, ci, c_dbl_p]),
        "gphocsStoreSync": (ci, [vp]),
        "gphocsStoreSetDebug": (ci, [vp, ci]),
        "gphocsStoreCheckMirror": (ci, [vp]),
        "gphocsKernelLaunchCount": (C.c_longlong, []),
        "gphocsCopyDeviceAsync": (ci, [vp, vp, C.c_longlong, vp]),
        "gphocsSetHostThreads": (ci, [ci]),
        "gphocsFiberSelfTest": (C.c_longlong, [ci, ci, ci, ci]),
        "gphocsHostAlloc": (vp, [C.c_longlong]),
        "gphocsHostFree": (ci, [vp]),
        # C. genealogy likelihood
        "gphocsGenCreate": (vp, [ci, ci, ci, ci, ci, c_int_p, c_int_p, c_int_p, c_int_p]),
        "gphocsGenDestroy": (ci, [vp]),
        "gphocsGenSetStream": (ci, [vp, vp]),
        "gphocsGenSetParams": (ci, [vp, c_dbl_p, c_dbl_p]),
        "gphocsGenSetEvents": (ci, [vp, c_ll_p, c_int_p, c_int_p, c_int_p, c_dbl_p]),
        "gphocsGenSetEventsPacked": (ci, [vp, c_int_p, C.POINTER(C.c_ushort), C.POINTER(C.c_ushort), c_dbl_p]),
        "gphocsGenEvaluate": (ci, [vp, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_ll_p, c_dbl_p, c_ll_p, c_dbl_p]),
        "gphocsGenEvaluateDevice": (ci, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "gphocsGenRecalc": (ci, [vp, ci, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p]),
        "gphocsGenRecalcAsync": (ci, [vp, ci, c_int_p, c_int_p, c_int_p, c_dbl_p, C.POINTER(vp)]),
        "gphocsGenGetStats": (ci, [vp, c_dbl_p, c_int_p, c_dbl_p, c_int_p]),
        "gphocsGenGetLineages": (ci, [vp, c_int_p]),
        "gphocsGenSync": (ci, [vp]),
        # D. device-resident MCMC steps
        "gphocsSamplerCreate": (vp, [vp, ci, ci, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p,
                                     c_dbl_p, c_dbl_p, c_int_p, C.c_ulonglong]),
        "gphocsSamplerDestroy": (ci, [vp]),
        "gphocsSamplerSetFinetunes": (ci, [vp, cd, cd, cd, cd]),
        "gphocsSamplerSetMigration": (ci, [vp, ci, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, c_int_p, c_int_p, c_dbl_p]),
        "gphocsSamplerSetMigFinetunes": (ci, [vp, cd, cd]),
        "gphocsSamplerSetAncient": (ci, [vp, c_int_p, c_dbl_p, cd, cd]),
        "gphocsSamplerSetAllReduce": (ci, [vp, vp, vp, C.c_longlong]),
        "gphocsNcclUniqueId": (ci, [C.c_char_p]),
        "gphocsSamplerInitNccl": (ci, [vp, C.c_char_p, ci, ci, C.c_longlong]),
        "gphocsSamplerIterate": (ci, [vp, ci, c_dbl_p]),
        "gphocsSamplerTraceWidth": (ci, [vp]),
        "gphocsSamplerOpenTrace": (ci, [vp, C.c_char_p, C.POINTER(C.c_char_p), cd, cd, ci]),
        "gphocsSamplerCloseTrace": (ci, [vp]),
        "gphocsSamplerSetStepwise": (ci, [vp, ci]),
        "gphocsSamplerSetMigRates": (ci, [vp, c_dbl_p]),
        "gphocsSamplerGetState": (ci, [vp, c_dbl_p, c_dbl_p, c_ll_p, c_ll_p]),
        "gphocsSamplerEvalCounters": (ci, [vp, C.POINTER(C.c_ulonglong), ci]),
        "gphocsSamplerCheck": (ci, [vp, c_dbl_p, c_dbl_p]),
        "gphocsSamplerDownload": (ci, [vp, c_int_p]),
        "gphocsSamplerGetStats": (ci, [vp, c_dbl_p, c_int_p, c_dbl_p, c_int_p]),
        # E. alignment ingest
        "gphocsReadSeqFile": (vp, [C.c_char_p, ci, C.POINTER(C.c_char_p), ci, ci]),
        "gphocsPhasePatterns": (vp, [ci, ci, C.c_char_p, c_int_p, C.c_char_p, c_int_p, ci, ci]),
        "gphocsAlignmentDims": (ci, [vp, c_int_p, c_int_p, c_int_p, c_int_p]),
        "gphocsAlignmentGet": (ci, [vp, c_ll_p, c_ll_p, C.c_char_p, c_int_p, c_int_p, C.c_char_p]),
        "gphocsAlignmentLocusName": (C.c_char_p, [vp, ci]),
        "gphocsAlignmentTimings": (ci, [vp, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_ll_p, c_ll_p]),
        "gphocsAlignmentFree": (None, [vp]),
        "gphocsStoreFromAlignment": (vp, [vp, ci]),
        "readSeqFile": (ci, [C.c_char_p, ci, C.POINTER(C.c_char_p), ci]),
        "processHetPatterns": (ci, [C.POINTER(C.c_char_p), c_int_p, ci, cus, vp, vp, c_int_p]),
        "freeAlignmentData": (ci, []),
        "printAlignmentError": (None, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dbl_p)


def _lp(a):
    return None if a is None else a.ctypes.data_as(c_ll_p)


class LociStore:
    """Batched engine (GphocsStore): all loci resident in HBM."""

    def __init__(self, n, patt_start, unph_start, chars, num_phases, counts, device=0, stream=None):
        self.lib = lib()
        self.n, self.N = int(n), 2 * int(n) - 1
        self.L = len(patt_start) - 1
        ps = np.ascontiguousarray(patt_start, np.int64)
        us = np.ascontiguousarray(unph_start, np.int64)
        ch = np.ascontiguousarray(chars, np.uint8)
        self.h = self.lib.gphocsStoreCreate(device, self.L, self.n, _lp(ps), _lp(us), ch.ctypes.data_as(C.c_char_p),
                                            _ip(_i32(num_phases)), _ip(_i32(counts)))
        if not self.h:
            raise RuntimeError("gphocsStoreCreate failed (no CUDA device? malformed patterns?) — no CPU fallback")
        self.h = C.c_void_p(self.h)
        if stream is not None:
            self.lib.gphocsStoreSetStream(self.h, C.c_void_p(stream))

    @classmethod
    def _adopt(cls, handle, L, n):
        s = cls.__new__(cls)
        s.lib, s.n, s.N, s.L, s.h = lib(), int(n), 2 * int(n) - 1, int(L), C.c_void_p(handle)
        return s

    @classmethod
    def from_workload(cls, w, device=0, stream=None):
        s = cls(w.n, w.patt_start, w.unph_start, w.chars, w.num_phases, w.counts, device, stream)
        s.set_trees(w.father, w.left, w.right, w.age, w.root)
        s.set_rates(w.rate)
        return s

    def close(self):
        if self.h:
            self.lib.gphocsStoreDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc})")

    def set_trees(self, father, left, right, age, root, ids=None):
        f, l, r, a, ro = _i32(father), _i32(left), _i32(right), _f64(age), _i32(root)
        k = len(ro)
        self._check(self.lib.gphocsStoreSetTrees(self.h, k, _ip(None if ids is None else _i32(ids)), _ip(f), _ip(l),
                                                 _ip(r), _dp(a), _ip(ro)), "gphocsStoreSetTrees")

    def set_trees_packed(self, topo16, age, root):
        """Loci 0..k-1 in the device's wire format: topo16[k][N][3] int16 (father, left, right), see pack_trees."""
        t = np.ascontiguousarray(topo16, np.int16)
        a, ro = _f64(age), _i32(root)
        self._check(self.lib.gphocsStoreSetTreesPacked(self.h, len(ro), t.ctypes.data_as(C.POINTER(C.c_short)), _dp(a),
                                                       _ip(ro)), "gphocsStoreSetTreesPacked")

    def get_trees(self, ids=None):
        k = self.L if ids is None else len(ids)
        f, l, r = (np.zeros((k, self.N), np.int32) for _ in range(3))
        a = np.zeros((k, self.N))
        ro = np.zeros(k, np.int32)
        self._check(self.lib.gphocsStoreGetTrees(self.h, k, _ip(None if ids is None else _i32(ids)), _ip(f), _ip(l),
                                                 _ip(r), _dp(a), _ip(ro)), "gphocsStoreGetTrees")
        return f, l, r, a, ro

    def set_rates(self, rates, ids=None):
        r = _f64(rates)
        self._check(self.lib.gphocsStoreSetRates(self.h, len(r), _ip(None if ids is None else _i32(ids)), _dp(r)),
                    "gphocsStoreSetRates")

    def get_rates(self, ids=None):
        out = np.zeros(self.L if ids is None else len(ids))
        self._check(self.lib.gphocsStoreGetRates(self.h, len(out), _ip(None if ids is None else _i32(ids)), _dp(out)),
                    "gphocsStoreGetRates")
        return out

    def apply_ops_async(self, ops):
        """Edit records (sorted by locus, page-locked: pinned_like) for the device copy only; returns at once."""
        assert ops.dtype == OP_DTYPE and ops.flags["C_CONTIGUOUS"]
        self._check(self.lib.gphocsStoreApplyOpsAsync(self.h, len(ops), ops.ctypes.data_as(C.c_void_p)), "gphocsStoreApplyOpsAsync")

    def apply_ops(self, ops, want_status=False):
        ops = np.ascontiguousarray(ops, OP_DTYPE)
        st = np.zeros(len(ops), np.int32) if want_status else None
        self._check(self.lib.gphocsStoreApplyOps(self.h, len(ops), ops.ctypes.data_as(C.c_void_p), _ip(st)),
                    "gphocsStoreApplyOps")
        return st

    def evaluate(self, use_old, ids=None, want_sum=False, out=None):
        k = self.L if ids is None else len(ids)
        out = np.zeros(k) if out is None else out
        s = C.c_double(0.0)
        self._check(self.lib.gphocsStoreEvaluate(self.h, k, _ip(None if ids is None else _i32(ids)), int(use_old), _dp(out),
                                                 C.byref(s) if want_sum else None), "gphocsStoreEvaluate")
        return (out, s.value) if want_sum else out

    def evaluate_device(self, use_old):
        """Launch only; returns (device pointer of lnL[L], device pointer of the summed lnL)."""
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.lib.gphocsStoreEvaluateDevice(self.h, int(use_old), C.byref(a), C.byref(b)),
                    "gphocsStoreEvaluateDevice")
        return a.value, b.value

    def lnl(self, ids=None):
        k = self.L if ids is None else len(ids)
        out = np.zeros(k)
        self._check(self.lib.gphocsStoreGetLnL(self.h, k, _ip(None if ids is None else _i32(ids)), _dp(out)), "gphocsStoreGetLnL")
        return out

    def clv(self, locus, node, saved=False, P=None):
        out = np.zeros(4 * P)
        self._check(self.lib.gphocsStoreGetClv(self.h, locus, node, int(saved), _dp(out)), "gphocsStoreGetClv")
        return out.reshape(P, 4)

    def sync(self):
        self._check(self.lib.gphocsStoreSync(self.h), "gphocsStoreSync")

    def set_debug(self, on=True):
        self.lib.gphocsStoreSetDebug(self.h, int(on))

    def check_mirror(self):
        return self.lib.gphocsStoreCheckMirror(self.h)

    @property
    def num_columns(self):
        return self.lib.gphocsStoreNumColumns(self.h)

    @property
    def device_bytes(self):
        return self.lib.gphocsStoreDeviceBytes(self.h)


def pinned_like(a):
    """Copy of array `a` in page-locked host memory (gphocsHostAlloc); keeps the allocation alive with the array."""
    a = np.ascontiguousarray(a)
    if a.nbytes == 0:
        return a
    L = lib()
    p = L.gphocsHostAlloc(a.nbytes)
    if not p:
        raise MemoryError("gphocsHostAlloc failed")
    buf = (C.c_char * a.nbytes).from_address(p)
    out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape)
    out[...] = a
    _PINNED.append((p, buf))
    return out


_PINNED = []


def free_pinned():
    L = lib()
    while _PINNED:
        p, _ = _PINNED.pop()
        L.gphocsHostFree(C.c_void_p(p))


def pack_trees(father, left, right):
    """int32 [L][N] topology arrays -> the wire format of gphocsStoreSetTreesPacked: int16 [L][N][3]."""
    return np.ascontiguousarray(np.stack([np.asarray(father), np.asarray(left), np.asarray(right)], axis=-1).astype(np.int16))


def pack_events(ev_start, pop_start, ev_type, ev_id):
    """The arrays of gphocsGenSetEvents -> (evStart32, popStart16, evCode16) of gphocsGenSetEventsPacked."""
    es = np.asarray(ev_start, np.int64)
    t = np.asarray(ev_type, np.int64)[es[0]:es[-1]]
    i = np.asarray(ev_id, np.int64)[es[0]:es[-1]]
    band = (t == 1) | (t == 3) | (t == 4)          # IN_MIG, MIG_BAND_START, MIG_BAND_END carry a band id
    code = (t | (np.where(band, i, 0) << 3)).astype(np.uint16)
    return (es - es[0]).astype(np.int32), np.asarray(pop_start).astype(np.uint16), code


def make_ops(locus, type_, a=0, b=0, x=0.0):
    """Edit-record array from broadcastable columns."""
    locus = np.atleast_1d(np.asarray(locus))
    ops = np.zeros(len(locus), OP_DTYPE)
    ops["locus"], ops["type"], ops["a"], ops["b"], ops["x"] = locus, type_, a, b, x
    return ops


class Genealogy:
    """Genealogy likelihood over flattened event chains (GphocsGenealogy)."""

    def __init__(self, L, pops, device=0, stream=None):
        self.lib = lib()
        self.L = int(L)
        self.Q = len(pops["father"])
        self.C = len(pops["samples_per_pop"])
        self.B = len(pops["band_src"])
        self.h = self.lib.gphocsGenCreate(device, self.L, self.Q, self.C, self.B, _ip(_i32(pops["father"])),
                                          _ip(_i32(pops["son0"])), _ip(_i32(pops["son1"])), _ip(_i32(pops["samples_per_pop"])))
        if not self.h:
            raise RuntimeError("gphocsGenCreate failed — no CPU fallback")
        self.h = C.c_void_p(self.h)
        if stream is not None:
            self.lib.gphocsGenSetStream(self.h, C.c_void_p(stream))
        self.set_params(pops["theta"], pops["band_rate"])
        self.E = 0

    def close(self):
        if self.h:
            self.lib.gphocsGenDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, theta, mig_rate):
        m = _f64(mig_rate) if self.B else np.zeros(1)
        if self.lib.gphocsGenSetParams(self.h, _dp(_f64(theta)), _dp(m)) != 0:
            raise RuntimeError("gphocsGenSetParams failed")

    def set_events(self, ev_start, pop_start, ev_type, ev_id, ev_time):
        es = np.ascontiguousarray(ev_start, np.int64)
        self.E = int(es[-1] - es[0])
        if self.lib.gphocsGenSetEvents(self.h, _lp(es), _ip(_i32(pop_start)), _ip(_i32(ev_type)), _ip(_i32(ev_id)),
                                       _dp(_f64(ev_time))) != 0:
            raise RuntimeError("gphocsGenSetEvents failed")

    def set_events_packed(self, ev_start32, pop_start16, ev_code16, ev_time):
        """The snapshot in the device's own format, see pack_events."""
        es = np.ascontiguousarray(ev_start32, np.int32)
        ps = np.ascontiguousarray(pop_start16, np.uint16)
        code = np.ascontiguousarray(ev_code16, np.uint16)
        self.E = int(es[-1])
        if self.lib.gphocsGenSetEventsPacked(self.h, _ip(es), ps.ctypes.data_as(C.POINTER(C.c_ushort)),
                                             code.ctypes.data_as(C.POINTER(C.c_ushort)), _dp(_f64(ev_time))) != 0:
            raise RuntimeError("gphocsGenSetEventsPacked failed")

    def evaluate(self, per_locus_stats=True):
        L, Q, B = self.L, self.Q, self.B
        lnl = np.zeros(L)
        coal = np.zeros((L, Q)) if per_locus_stats else None
        ncoal = np.zeros((L, Q), np.int32) if per_locus_stats else None
        mig = np.zeros((L, max(B, 1))) if per_locus_stats else None
        nmig = np.zeros((L, max(B, 1)), np.int32) if per_locus_stats else None
        if per_locus_stats and B:
            mig = np.zeros((L, B))
            nmig = np.zeros((L, B), np.int32)
        tc, tm = np.zeros(Q), np.zeros(max(B, 1))
        tnc, tnm = np.zeros(Q, np.int64), np.zeros(max(B, 1), np.int64)
        s = C.c_double()
        rc = self.lib.gphocsGenEvaluate(self.h, _dp(lnl), _dp(coal), _ip(ncoal), _dp(mig), _ip(nmig), _dp(tc), _lp(tnc),
                                        _dp(tm), _lp(tnm), C.byref(s))
        if rc != 0:
            raise RuntimeError("gphocsGenEvaluate failed")
        return dict(lnl=lnl, coal=coal, num_coals=ncoal, mig=None if mig is None else mig[:, :B],
                    num_migs=None if nmig is None else nmig[:, :B], total_coal=tc, total_num_coals=tnc,
                    total_mig=tm[:B], total_num_migs=tnm[:B], sum_lnl=s.value)

    def evaluate_device(self):
        a, b = C.c_void_p(), C.c_void_p()
        v = self.lib.gphocsGenEvaluateDevice(self.h, C.byref(a), C.byref(b))
        if v < 0:
            raise RuntimeError("gphocsGenEvaluateDevice failed")
        return a.value, b.value, v

    def recalc(self, locus, pop, times_start, times):
        """recalcStats for the listed (locus, population) chains with new elapsed times; returns its return values."""
        locus, pop, ts = _i32(locus), _i32(pop), _i32(times_start)
        t = np.ascontiguousarray(times, np.float64)
        out = np.zeros(len(locus))
        if self.lib.gphocsGenRecalc(self.h, len(locus), _ip(locus), _ip(pop), _ip(ts), _dp(t), _dp(out)) != 0:
            raise RuntimeError("gphocsGenRecalc failed")
        return out

    def recalc_async(self, locus, pop, times_start, times):
        """gphocsGenRecalcAsync on page-locked int32 / float64 arrays (pinned_like); returns the device pointer of the deltas."""
        out = C.c_void_p()
        if self.lib.gphocsGenRecalcAsync(self.h, len(locus), _ip(locus), _ip(pop), _ip(times_start), _dp(times), C.byref(out)) != 0:
            raise RuntimeError("gphocsGenRecalcAsync failed")
        return out.value

    def sync(self):
        if self.lib.gphocsGenSync(self.h) != 0:
            raise RuntimeError("gphocsGenSync reported refused chains or a CUDA error")

    def stats_only(self):
        """Per-locus statistics as stored on the device, without evaluating."""
        L, Q, B = self.L, self.Q, self.B
        coal, ncoal = np.zeros((L, Q)), np.zeros((L, Q), np.int32)
        mig, nmig = np.zeros((L, max(B, 1))), np.zeros((L, max(B, 1)), np.int32)
        if B:
            mig, nmig = np.zeros((L, B)), np.zeros((L, B), np.int32)
        if self.lib.gphocsGenGetStats(self.h, _dp(coal), _ip(ncoal), _dp(mig), _ip(nmig)) != 0:
            raise RuntimeError("gphocsGenGetStats failed")
        return dict(coal=coal, num_coals=ncoal, mig=mig[:, :B], num_migs=nmig[:, :B])

    def lineages(self):
        out = np.zeros(self.E, np.int32)
        if self.lib.gphocsGenGetLineages(self.h, _ip(out)) != 0:
            raise RuntimeError("gphocsGenGetLineages failed")
        return out


class ScalarLocus:
    """One locus driven through the reference's own LocusData entry points exported by the library
    (createLocusData ... getNodeSon) — the un-modified-caller path of INTEGRATION.md §1."""

    def __init__(self, n, chars, num_phases, counts, rate=1.0):
        self.lib = lib()
        self.n, self.N, self.P = n, 2 * n - 1, len(num_phases)
        self.h = C.c_void_p(self.lib.createLocusData(n, 1))
        chars = np.ascontiguousarray(chars, np.uint8)
        rows = (C.c_char_p * max(self.P, 1))(*[chars[p].tobytes() for p in range(self.P)])
        ph, ct = _i32(num_phases), _i32(counts)
        if self.lib.initializeLocusData(self.h, rows, self.P, _ip(ph), _ip(ct)) != 0:
            self.lib.freeLocusData(self.h)
            raise ValueError("initializeLocusData failed")
        self.lib.setLocusMutationRate(self.h, rate)

    def set_tree(self, father, left, right, age, root):
        f, l, r, a = _i32(father).copy(), _i32(left).copy(), _i32(right).copy(), _f64(age).copy()
        t = GenericBinaryTree(self.n, int(root), None, _ip(f), _ip(l), _ip(r), _dp(a), None)
        self.lib.copyGenericTreeToLocus(self.h, C.byref(t))

    def tree(self):
        N = self.N
        f = np.array([self.lib.getNodeFather(self.h, i) for i in range(N)], np.int32)
        l = np.array([self.lib.getNodeSon(self.h, i, 0) for i in range(N)], np.int32)
        r = np.array([self.lib.getNodeSon(self.h, i, 1) for i in range(N)], np.int32)
        a = np.array([self.lib.getNodeAge(self.h, i) for i in range(N)])
        return f, l, r, a, self.lib.getLocusRoot(self.h)

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

    def free(self):
        if self.h:
            self.lib.freeLocusData(self.h)
            self.h = None


class Alignment:
    """Sequence file -> initializeLocusData's arguments for every locus (GphocsAlignment, header group E)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("the sequence file was rejected (message on stderr)")
        self.lib = lib()
        self.h = C.c_void_p(handle)
        d = [C.c_int() for _ in range(4)]
        self.lib.gphocsAlignmentDims(self.h, *[C.byref(x) for x in d])
        self.L, self.n, self.P, self.U = (x.value for x in d)
        self.patt_start, self.unph_start = np.zeros(self.L + 1, np.int64), np.zeros(self.L + 1, np.int64)
        self.chars = np.zeros((max(self.P, 1), self.n), np.uint8)
        self.canon = np.zeros((max(self.U, 1), self.n), np.uint8)
        self.num_phases, self.counts = np.zeros(max(self.P, 1), np.int32), np.zeros(max(self.U, 1), np.int32)
        self.lib.gphocsAlignmentGet(self.h, _lp(self.patt_start), _lp(self.unph_start), self.chars.ctypes.data_as(C.c_char_p),
                                    _ip(self.num_phases), _ip(self.counts), self.canon.ctypes.data_as(C.c_char_p))
        self.chars, self.canon = self.chars[:self.P], self.canon[:self.U]
        self.num_phases, self.counts = self.num_phases[:self.P], self.counts[:self.U]

    @classmethod
    def read(cls, path, sample_names, num_loci_to_read=0, device=0):
        """sample_names: one entry per haploid slot, "" for the second slot of a diploid sample."""
        arr = (C.c_char_p * len(sample_names))(*[nm.encode() for nm in sample_names])
        return cls(lib().gphocsReadSeqFile(str(path).encode(), len(sample_names), arr, int(num_loci_to_read), int(device)))

    @classmethod
    def phase(cls, start, patterns, counts, is_diploid, break_symmetries=1, device=0):
        """processHetPatterns alone: canonical patterns uint8 [sum U][n] of several loci (CSR offsets `start`)."""
        patterns = np.ascontiguousarray(patterns, np.uint8)
        dip = bytes(1 if d else 0 for d in is_diploid)
        return cls(lib().gphocsPhasePatterns(len(start) - 1, len(dip), dip, _ip(_i32(start)), patterns.ctypes.data_as(C.c_char_p),
                                             _ip(_i32(counts)), int(break_symmetries), int(device)))

    def locus(self, l):
        """(chars [P][n], num_phases [P], counts [U]) of one locus"""
        a, b, c, d = self.patt_start[l], self.patt_start[l + 1], self.unph_start[l], self.unph_start[l + 1]
        return self.chars[a:b], self.num_phases[a:b], self.counts[c:d]

    def locus_name(self, l):
        return self.lib.gphocsAlignmentLocusName(self.h, int(l)).decode()

    def timings(self):
        t = [C.c_double() for _ in range(4)]
        b = [C.c_longlong() for _ in range(2)]
        self.lib.gphocsAlignmentTimings(self.h, *[C.byref(x) for x in t], *[C.byref(x) for x in b])
        return dict(parse_s=t[0].value, h2d_s=t[1].value, kernel_s=t[2].value, d2h_s=t[3].value, raw_bytes=b[0].value,
                    out_bytes=b[1].value)

    def store(self, device=0):
        h = self.lib.gphocsStoreFromAlignment(self.h, int(device))
        if not h:
            raise RuntimeError("gphocsStoreFromAlignment failed")
        return LociStore._adopt(h, self.L, self.n)

    def close(self):
        if self.h:
            self.lib.gphocsAlignmentFree(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sampler:
    """Device-resident MCMC update steps (GphocsSampler) over the loci of a LociStore."""
    MOVES = ("coal_time", "spr", "theta", "tau", "mixing", "mig_rate", "mig_time", "tau_conflicts", "locus_rate",
             "sample_age")
    MAX_MIGS = 10

    def __init__(self, store, pops, node_pop, theta_prior=(1.0, 1000.0), tau_prior=None, seed=1, finetunes=None,
                 migration=None, mig_prior=(0.002, 0.00001), mig_finetunes=None, estimate_sample_age=None,
                 locus_rate_finetune=0.0, rate_alpha=1.0):
        """pops["sample_age"] (ages of the current populations' samples) must agree with the leaf ages in the store;
        estimate_sample_age[C] marks the ones that are parameters; locus_rate_finetune > 0 turns on locus-rate moves.
        migration = (mig_start[L+1], mig_branch, mig_band, mig_age) CSR arrays of the genealogies' migration events
        (synth.Workload fields); bands and their rates come from pops["band_src"/"band_tgt"/"band_rate"]."""
        self.lib = lib()
        self.store = store
        self.Q = len(pops["father"])
        self.C = len(pops["samples_per_pop"])
        Q = self.Q
        ta = np.full(Q, float(theta_prior[0]))
        tb = np.full(Q, float(theta_prior[1]))
        if tau_prior is None:        # the control files written by synth.write_control_file: alpha 1, beta 1/tau-initial
            aa = np.full(Q, 1.0)
            ab = np.array([1.0 / t if t > 0 else 1.0 for t in pops["age"]])
            aa[:self.C] = ab[:self.C] = 0.0      # sample ages: no prior density (PopulationTree.c:121)
        else:
            aa, ab = (np.ascontiguousarray(x, np.float64) for x in tau_prior)
        tau0 = _f64(pops["age"]).copy()
        if "sample_age" in pops:
            tau0[:self.C] = pops["sample_age"]
        self.h = self.lib.gphocsSamplerCreate(store.h, Q, self.C, _ip(_i32(pops["father"])), _ip(_i32(pops["son0"])),
                                              _ip(_i32(pops["son1"])), _ip(_i32(pops["samples_per_pop"])),
                                              _dp(_f64(pops["theta"])), _dp(tau0), _dp(ta), _dp(tb), _dp(aa), _dp(ab),
                                              _ip(_i32(node_pop)), int(seed))
        if not self.h:
            raise RuntimeError("gphocsSamplerCreate failed")
        self.h = C.c_void_p(self.h)
        self.B = 0
        if migration is not None and len(pops["band_src"]) > 0:
            ms, mbr, mbd, mag = migration
            L, B = store.L, len(pops["band_src"])
            num = np.diff(np.asarray(ms)).astype(np.int32)
            br = np.full((L, self.MAX_MIGS), -1, np.int32)
            bd = np.zeros((L, self.MAX_MIGS), np.int32)
            ag = np.zeros((L, self.MAX_MIGS))
            for l in np.nonzero(num)[0]:
                a, b = int(ms[l]), int(ms[l + 1])
                br[l, :b - a], bd[l, :b - a], ag[l, :b - a] = mbr[a:b], mbd[a:b], mag[a:b]
            rc = self.lib.gphocsSamplerSetMigration(self.h, B, _ip(_i32(pops["band_src"])), _ip(_i32(pops["band_tgt"])),
                                                    _dp(_f64(pops["band_rate"])), _dp(np.full(B, float(mig_prior[0]))),
                                                    _dp(np.full(B, float(mig_prior[1]))), _ip(num), _ip(br), _ip(bd), _dp(ag))
            if rc != 0:
                raise RuntimeError("gphocsSamplerSetMigration failed")
            self.B = B
            if mig_finetunes is not None:
                self.lib.gphocsSamplerSetMigFinetunes(self.h, float(mig_finetunes[0]), float(mig_finetunes[1]))
        if estimate_sample_age is not None or locus_rate_finetune > 0.0:
            est = _i32(estimate_sample_age if estimate_sample_age is not None else np.zeros(self.C))
            self.lib.gphocsSamplerSetAncient(self.h, _ip(est), None, float(locus_rate_finetune), float(rate_alpha))
        self.width = self.lib.gphocsSamplerTraceWidth(self.h)
        if finetunes is not None:
            self.lib.gphocsSamplerSetFinetunes(self.h, *[float(x) for x in finetunes])

    def set_all_reduce(self, fn, locus_offset=0):
        """fn(numpy float64 vector) must sum it over all ranks in place (e.g. through an NCCL all-reduce)."""
        proto = C.CFUNCTYPE(C.c_int, c_dbl_p, C.c_int, C.c_void_p)

        def thunk(buf, count, _ctx):
            try:
                fn(np.ctypeslib.as_array(buf, shape=(count,)))
                return 0
            except Exception as e:  # noqa: BLE001 - reported through the C return code
                print(f"all-reduce hook failed: {e}")
                return -1
        self._ar = proto(thunk)      # keep the callback alive
        self.lib.gphocsSamplerSetAllReduce(self.h, C.cast(self._ar, C.c_void_p), None, int(locus_offset))

    def open_trace(self, path, pop_names, theta_tau_print=10000.0, mig_rate_print=0.001, sample_skip=0):
        """Trace file in the reference's format; rows are appended by iterate()."""
        arr = (C.c_char_p * len(pop_names))(*[x.encode() for x in pop_names])
        if self.lib.gphocsSamplerOpenTrace(self.h, str(path).encode(), arr, float(theta_tau_print), float(mig_rate_print),
                                           int(sample_skip)) != 0:
            raise RuntimeError("gphocsSamplerOpenTrace failed")

    def eval_counters(self, reset=False):
        """(incremental locus evaluations, their algorithmic bytes per SURVEY.md 8d) since the last reset; the first
        call switches the accounting on."""
        out = (C.c_ulonglong * 2)()
        if self.lib.gphocsSamplerEvalCounters(self.h, out, int(bool(reset))) != 0:
            raise RuntimeError("gphocsSamplerEvalCounters failed")
        return int(out[0]), int(out[1])

    def set_mig_rates(self, rates):
        r = np.ascontiguousarray(rates, np.float64)
        if self.lib.gphocsSamplerSetMigRates(self.h, _dp(r)) != 0:
            raise RuntimeError("gphocsSamplerSetMigRates failed")

    def set_stepwise(self, on):
        """1: per-node launches even where the one-launch sweep applies (same chain either way)."""
        self.lib.gphocsSamplerSetStepwise(self.h, int(bool(on)))

    def close_trace(self):
        self.lib.gphocsSamplerCloseTrace(self.h)

    def init_nccl(self, rank, world, locus_offset=0, broadcast=None):
        """Own NCCL communicator of the library over `world` ranks.  broadcast(bytes_or_None) -> bytes must hand rank 0's
        128-byte id to every rank; default: torch.distributed.broadcast_object_list."""
        uid = C.create_string_buffer(128)
        if rank == 0 and self.lib.gphocsNcclUniqueId(uid) != 0:
            raise RuntimeError("gphocsNcclUniqueId failed")
        if broadcast is None:
            import torch.distributed as dist
            box = [uid.raw if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            raw = box[0]
        else:
            raw = broadcast(uid.raw if rank == 0 else None)
        if self.lib.gphocsSamplerInitNccl(self.h, raw, int(rank), int(world), int(locus_offset)) != 0:
            raise RuntimeError("gphocsSamplerInitNccl failed")

    def iterate(self, iterations, trace=True):
        out = np.zeros((iterations, self.width)) if trace else None
        if self.lib.gphocsSamplerIterate(self.h, int(iterations), _dp(out)) != 0:
            raise RuntimeError("gphocsSamplerIterate failed")
        return out

    def state(self):
        th, ta = np.zeros(self.Q), np.zeros(self.Q)
        acc, prop = np.zeros(10, np.int64), np.zeros(10, np.int64)
        self.lib.gphocsSamplerGetState(self.h, _dp(th), _dp(ta), _lp(acc), _lp(prop))
        return dict(theta=th, tau=ta, accepted=dict(zip(self.MOVES, acc)), proposed=dict(zip(self.MOVES, prop)))

    def check(self):
        a, b = C.c_double(), C.c_double()
        v = self.lib.gphocsSamplerCheck(self.h, C.byref(a), C.byref(b))
        return v, a.value, b.value

    def stats(self):
        L, Q, B = self.store.L, self.Q, max(self.B, 1)
        coal, mig = np.zeros((L, Q)), np.zeros((L, B))
        nc, nm = np.zeros((L, Q), np.int32), np.zeros((L, B), np.int32)
        self.lib.gphocsSamplerGetStats(self.h, _dp(coal), _ip(nc), _dp(mig), _ip(nm))
        return dict(coal=coal, num_coals=nc, mig=mig[:, :self.B], num_migs=nm[:, :self.B])

    def download(self):
        out = np.zeros((self.store.L, self.store.N), np.int32)
        if self.lib.gphocsSamplerDownload(self.h, _ip(out)) != 0:
            raise RuntimeError("gphocsSamplerDownload failed")
        return out

    def close(self):
        if self.h:
            self.lib.gphocsSamplerDestroy(self.h)
            self.h = None
