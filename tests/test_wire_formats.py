"""CPU: the host-side helpers that write the library's wire formats (gphocsStoreSetTreesPacked /
gphocsGenSetEventsPacked, include/gphocs_b200.h) against the rule the device applies to int32 input (k_gen_pack in
gen_kernels.cuh: evCode = type | band << 3, band only for IN_MIG / MIG_BAND_START / MIG_BAND_END)."""
import importlib

import numpy as np

gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")


def test_pack_trees_layout():
    w = synth.generate(synth.config("dip8mig"), 50, seed=5)
    t = gp.pack_trees(w.father, w.left, w.right)
    N = 2 * w.n - 1
    assert t.dtype == np.int16 and t.shape == (50, N, 3) and t.flags["C_CONTIGUOUS"]
    assert np.array_equal(t[..., 0], w.father) and np.array_equal(t[..., 1], w.left) and np.array_equal(t[..., 2], w.right)
    assert t.min() >= -1 and t.max() < N


def test_pack_events_codes_and_offsets():
    w = synth.generate(synth.config("pop6mig4"), 200, seed=6)
    es, ps, code = gp.pack_events(w.ev_start, w.pop_start, w.ev_type, w.ev_id)
    assert es.dtype == np.int32 and ps.dtype == np.uint16 and code.dtype == np.uint16
    assert es[0] == 0 and es[-1] == len(code) == w.ev_start[-1] - w.ev_start[0]
    assert np.array_equal(np.diff(es), np.diff(w.ev_start))
    assert np.array_equal(ps, w.pop_start)
    # chains of a locus are stored in population order and cover all its events
    assert np.all(ps[:, 0] == 0) and np.array_equal(ps[:, -1], np.diff(es)) and np.all(np.diff(ps.astype(np.int64), axis=1) >= 0)
    t = np.asarray(w.ev_type)[w.ev_start[0]:w.ev_start[-1]]
    i = np.asarray(w.ev_id)[w.ev_start[0]:w.ev_start[-1]]
    assert np.array_equal(code & 7, t)
    band = (t == 1) | (t == 3) | (t == 4)
    assert np.array_equal((code >> 3)[band], i[band]) and np.all((code >> 3)[~band] == 0)
    assert band.any() and (code >> 3).max() < len(w.pops["band_src"])
    # a snapshot that does not start at event 0 is rebased
    es2, _, code2 = gp.pack_events(w.ev_start + 7, w.pop_start, np.concatenate([np.zeros(7, np.int32), w.ev_type]),
                                   np.concatenate([np.zeros(7, np.int32), w.ev_id]))
    assert np.array_equal(es2, es) and np.array_equal(code2, code)
