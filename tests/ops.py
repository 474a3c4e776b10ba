"""Random, valid proposal sequences over the LocusData call surface (age moves, SPR, rescaling,
ancient-sample moves, rate changes, accept/reject) — the traffic the reference's MCMC update steps
generate (GPhoCS.c:2287-4916).  Any backend exposing
    tree() compute(use_old) lnl() adjust_age() spr() scale_all() revert() reset() set_rate()
can be driven; two backends fed the same seed must produce identical traces.
"""
import numpy as np


def random_tree(n, rng, scale=1e-3, leaf_ages=None):
    """Random binary genealogy with leaves 0..n-1 and internal nodes n..2n-2 created in age order."""
    N = 2 * n - 1
    father = np.full(N, -1, np.int32)
    left = np.full(N, -1, np.int32)
    right = np.full(N, -1, np.int32)
    age = np.zeros(N)
    if leaf_ages is not None:
        age[:n] = leaf_ages
    active = list(range(n))
    t = float(age[:n].max())
    for node in range(n, N):
        k = len(active)
        t += rng.exponential(scale / (k * (k - 1) / 2))
        i, j = rng.choice(k, 2, replace=False)
        a, b = active[i], active[j]
        left[node], right[node] = a, b
        father[a] = father[b] = node
        age[node] = t
        active = [x for x in active if x not in (a, b)] + [node]
    return father, left, right, age, N - 1


def random_patterns(n, P, rng, diploid_pairs=0, missing=0.0):
    """Random phased pattern block: chars [P,n] uint8, num_phases [P], counts [U]."""
    rows, phases, counts = [], [], []
    while len(rows) < P:
        base = rng.integers(0, 4, n)
        if rng.random() < 0.6:
            base[:] = base[0]
            k = rng.integers(0, n)
            base[k] = (base[k] + 1 + rng.integers(0, 3)) % 4
        col = np.array([ord("TCAG"[b]) for b in base], np.uint8)
        if missing > 0:
            col[rng.random(n) < missing] = ord("N")
        h = 0
        if diploid_pairs and rng.random() < 0.5:
            h = int(rng.integers(1, min(diploid_pairs, 3) + 1))
        group = 1 << h
        if len(rows) + group > P:
            h, group = 0, 1
        pairs = rng.choice(max(diploid_pairs, 1), h, replace=False) if h else []
        for mask in range(group):
            c = col.copy()
            for j, pr in enumerate(pairs):
                a, b = 2 * pr, 2 * pr + 1
                if c[a] == c[b]:
                    c[b] = ord("TCAG"[("TCAG".index(chr(c[a])) + 1) % 4]) if c[a] != ord("N") else c[b]
                if mask >> j & 1:
                    c[a], c[b] = c[b], c[a]
            rows.append(c)
            phases.append(group if mask == 0 else 0)
        counts.append(int(rng.integers(1, 400)))
    return np.array(rows, np.uint8), np.array(phases, np.int32), np.array(counts, np.int32)


def _subtree(left, right, s):
    out, stack = set(), [s]
    while stack:
        v = stack.pop()
        out.add(v)
        if left[v] >= 0:
            stack += [left[v], right[v]]
    return out


def propose(locus, rng, n, allow_leaf_age=False, force_kind=None, exclude=()):
    """Applies one random valid proposal to `locus`; returns a short description tuple.
    A node may be saved at most once per proposal round (LocusDataLikelihood.c:1860-1861), hence
    `exclude` for multi-node rounds."""
    father, left, right, age, root = locus.tree()
    N = 2 * n - 1
    kind = force_kind or rng.choice(["age", "age", "spr", "spr", "spr", "scale", "leaf"] if allow_leaf_age
                                    else ["age", "age", "spr", "spr", "spr", "scale"])
    if kind == "age":
        i = int(rng.integers(n, N))
        if i in exclude:
            return ("none",)
        lo = max(age[left[i]], age[right[i]])
        hi = age[father[i]] if father[i] >= 0 else age[i] * 1.3 + 1e-5
        new = lo + (hi - lo) * rng.random()
        locus.adjust_age(i, new)
        return ("age", i, new)
    if kind == "leaf":
        i = int(rng.integers(0, n))
        hi = age[father[i]]
        new = age[i] + (hi - age[i]) * 0.3 * rng.random()
        locus.adjust_age(i, new)
        return ("leaf", i, new)
    if kind == "scale":
        f = 0.9 + 0.2 * rng.random()
        d = locus.scale_all(f)
        return ("scale", f, d)
    # SPR
    for _ in range(50):
        s = int(rng.integers(0, N))
        if s == root:
            continue
        F = father[s]
        G = father[F]
        sib = left[F] + right[F] - s
        banned = _subtree(left, right, s)
        t = int(rng.integers(0, N))
        if t in banned:
            continue
        if t == sib:
            lo, hi = max(age[s], age[sib]), (age[G] if G >= 0 else age[F] * 1.3 + 1e-5)
        elif t == F:
            lo, hi = max(age[s], age[sib]), (age[G] if G >= 0 else age[F] * 1.3 + 1e-5)
        else:
            lo = max(age[s], age[t])
            tf = father[t]
            if tf == F:      # t is the sibling (handled above)
                continue
            hi = age[tf] if tf >= 0 else lo * 1.3 + 1e-5
            if t == root:
                hi = max(lo, age[t]) * 1.3 + 1e-5
        if hi <= lo:
            continue
        new = lo + (hi - lo) * (0.05 + 0.9 * rng.random())
        rc = locus.spr(s, t, new)
        return ("spr", s, t, new, rc)
    return ("none",)


def run_ops(locus, n, seed, steps, allow_leaf_age=False, rate_moves=False):
    """Drives `steps` proposal/evaluate/accept-or-reject rounds; returns the observable trace."""
    rng = np.random.default_rng(seed)
    trace = []
    trace.append(("init", locus.compute(0)))
    locus.reset()
    for _ in range(steps):
        if rate_moves and rng.random() < 0.1:
            locus.set_rate(0.5 + rng.random())
            trace.append(("rate", locus.compute(0)))
            if rng.random() < 0.5:
                locus.reset()
            else:
                locus.revert()   # NB: the reference restores the rate separately (GPhoCS.c:4667-4670)
            continue
        desc = []
        if rng.random() < 0.8:
            desc.append(propose(locus, rng, n, allow_leaf_age))
        else:   # several node ages moved before one evaluation, as UpdateTau's rubber band does
            moved = []
            for _k in range(int(rng.integers(2, 5))):
                d = propose(locus, rng, n, force_kind="age", exclude=moved)
                desc.append(d)
                if d[0] == "age":
                    moved.append(d[1])
        if desc[-1][0] == "scale":
            new = locus.lnl()
        else:
            new = locus.compute(1)
        accept = rng.random() < 0.5
        if accept:
            locus.reset()
        else:
            locus.revert()
        f, l, r, a, root = locus.tree()
        trace.append((tuple(desc), new, accept, locus.lnl(), f.copy(), l.copy(), r.copy(), a.copy(), root))
    trace.append(("final-full", locus.compute(0)))
    locus.reset()
    return trace


def traces_equal(t1, t2, rtol=0.0):
    """Compares two traces: integers/trees exactly, floats to rtol (0 = bit-exact)."""
    assert len(t1) == len(t2)
    for a, b in zip(t1, t2):
        assert len(a) == len(b), (a, b)
        for x, y in zip(a, b):
            _cmp(x, y, rtol)
    return True


def _cmp(x, y, rtol):
    if isinstance(x, tuple):
        assert len(x) == len(y), (x, y)
        for u, v in zip(x, y):
            _cmp(u, v, rtol)
    elif isinstance(x, np.ndarray):
        if x.dtype.kind == "f":
            assert np.allclose(x, y, rtol=rtol, atol=0.0) if rtol else np.array_equal(x, y), (x, y)
        else:
            assert np.array_equal(x, y), (x, y)
    elif isinstance(x, float):
        if rtol:
            assert abs(x - y) <= rtol * max(abs(x), abs(y), 1e-300), (x, y)
        else:
            assert x == y, (x, y)
    else:
        assert x == y, (x, y)
