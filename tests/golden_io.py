"""Loader for tests/golden/ref_*.npz (vectors dumped from the reference by make_golden.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["sample", "hap16", "dip8mig", "pop6mig4", "ancient"]


def load(name):
    z = np.load(os.path.join(HERE, "golden", f"ref_{name}.npz"))
    return {k: z[k] for k in z.files}


def pops_of(g):
    """Model.arrays()-style dict from a golden fixture."""
    return dict(theta=g["theta"], age=g["pop_age"], sample_age=g["sample_age"], father=g["pop_father"],
                son0=g["pop_son0"], son1=g["pop_son1"], samples_per_pop=g["samples_per_pop"],
                band_src=g["band_src"], band_tgt=g["band_tgt"], band_rate=g["band_rate"])


def locus_slices(g, l):
    p0, p1 = int(g["patt_start"][l]), int(g["patt_start"][l + 1])
    u0, u1 = int(g["unph_start"][l]), int(g["unph_start"][l + 1])
    e0, e1 = int(g["ev_start"][l]), int(g["ev_start"][l + 1])
    m0, m1 = int(g["mig_start"][l]), int(g["mig_start"][l + 1])
    return (p0, p1), (u0, u1), (e0, e1), (m0, m1)
