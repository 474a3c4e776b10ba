"""CPU, only where oracle/_ref was built (this container; it also travels to the GPU box):
oracle and compiled reference driven in lock-step through random proposal sequences."""
import numpy as np
import pytest

import ops
from oracle import bindings as ob

pytestmark = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.mark.parametrize("seed", range(24))
def test_random_proposal_sequences_bit_exact(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 20))
    P = int(rng.integers(1, 30))
    chars, ph, cnt = ops.random_patterns(n, P, rng, diploid_pairs=n // 2 if seed % 2 else 0,
                                         missing=0.1 if seed % 3 == 0 else 0.0)
    leaf_ages = np.where(rng.random(n) < 0.3, 1e-4, 0.0) if seed % 4 == 0 else None
    f, l, r, a, root = ops.random_tree(n, rng, leaf_ages=leaf_ages)
    rate = 0.5 + rng.random()
    tr = []
    for cls in (ob.RefLocus, ob.OracleLocus):
        lc = cls(n, chars, ph, cnt, rate)
        lc.set_tree(f, l, r, a, root)
        tr.append(ops.run_ops(lc, n, seed, 50, allow_leaf_age=(seed % 4 == 0), rate_moves=(seed % 5 == 0)))
    ops.traces_equal(tr[0], tr[1], rtol=0.0)


def test_check_locus_data_likelihood():
    rng = np.random.default_rng(5)
    n = 9
    chars, ph, cnt = ops.random_patterns(n, 11, rng)
    f, l, r, a, root = ops.random_tree(n, rng)
    for cls in (ob.RefLocus, ob.OracleLocus):
        lc = cls(n, chars, ph, cnt, 1.0)
        lc.set_tree(f, l, r, a, root)
        lc.compute(0)
        lc.reset()
        assert lc.check() == 1
