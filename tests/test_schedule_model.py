"""CPU: a Python model of the schedule k_eval builds for its column walk (g-phocs_b200/csrc/clv_kernels.cuh, phases
B-E) on random genealogies and random dirty sets.

It restates the rules of phases C-D2 — subtree weights, heavier child first, position = start + size - 1, where each
child's vector comes from (leaf mask / register top / stack slot / HBM record), which results are parked — and then
walks the schedule the way a column thread does, checking what the kernel relies on:
  * positions are a permutation of 0..k-1 and children precede parents,
  * a child marked TOP is the entry right before its parent,
  * a child marked STACK is found in the slot its parent expects, and no slot is overwritten while it is live,
  * a computed child marked GLOBAL (stack deeper than kStack) has been written before it is re-read,
  * parked results never exceed log2(number of leaves) levels.
The CUDA code itself is checked against the oracle on the GPU (tests/test_gpu_parity.py); this file guards the
algorithm the schedule encodes and documents it in executable form."""
import math
import random

K_STACK = 3   # kStack in clv_kernels.cuh


def random_tree(n, rng):
    N = 2 * n - 1
    father, left, right = [-1] * N, [-1] * N, [-1] * N
    active, ids = list(range(n)), list(range(n, N))
    rng.shuffle(ids)
    for v in ids:
        a = active.pop(rng.randrange(len(active)))
        b = active.pop(rng.randrange(len(active)))
        left[v], right[v], father[a], father[b] = a, b, v, v
        active.append(v)
    return father, left, right, active[0]


def build_schedule(n, father, left, right, root, need, k_stack=K_STACK):
    marked = [v for v in range(n, 2 * n - 1) if need[v]]
    size = [0] * (2 * n - 1)
    for v in marked:                                   # phase C: marked nodes per subtree
        a = v
        while a >= 0:
            size[a] += 1
            a = father[a]
    w = lambda x: size[x] if (x >= n and need[x]) else 0
    walk = {}
    for v in marked:                                   # phase D1: what v adds to the start of everything below it
        a, contrib = father[v], 0
        if a >= 0:
            l, r = left[a], right[a]
            first = l if w(l) >= w(r) else r
            if v != first:
                contrib = w(r) if v == l else w(l)
        walk[v] = (a, contrib)
    sched, max_depth = {}, 0
    for v in marked:                                   # phase D2
        start = depth = 0
        x = v
        while True:
            a, c = walk[x]
            start += c
            depth += c != 0
            if a < 0:
                break
            x = a
        l, r = left[v], right[v]
        left_first = w(l) >= w(r)
        A, B = (l, r) if left_first else (r, l)
        wA, wB = (w(l), w(r)) if left_first else (w(r), w(l))
        if A < n:
            kA = ("leaf", A)
        elif wA > 0 and wB == 0:
            kA = ("top", A)
        elif wA > 0 and depth < k_stack:
            kA = ("stack", depth, A)
        else:
            kA = ("global", A)
        kB = ("leaf", B) if B < n else ("top", B) if wB > 0 else ("global", B)
        push = None
        f = father[v]
        if f >= 0 and v != root:
            fl, fr = left[f], right[f]
            first = fl if w(fl) >= w(fr) else fr
            sibling = w(fr) if v == fl else w(fl)
            if v == first and sibling > 0:
                max_depth = max(max_depth, depth + 1)
                if depth < k_stack:
                    push = depth
        pos = start + size[v] - 1
        assert pos not in sched
        sched[pos] = (v, kA, kB, push)
    return sched, (size[root] if need[root] else 0), max_depth


def walk_schedule(n, left, right, need, sched, k, k_stack=K_STACK):
    top, stack, written, smem = None, [None] * k_stack, set(), 0
    for e in range(k):
        v, kA, kB, push = sched[e]
        for kind in (kA, kB):
            if kind[0] == "top":
                assert top == kind[1]
            elif kind[0] == "stack":
                assert stack[kind[1]] == kind[2]
                stack[kind[1]] = None                    # consumed
                smem += 1
            elif kind[0] == "global":
                assert not need[kind[1]] or kind[1] in written
            else:
                assert kind[1] < n
        assert {kA[-1], kB[-1]} == {left[v], right[v]}
        written.add(v)
        if push is not None:
            assert stack[push] is None                   # a live slot is never overwritten
            stack[push] = v
            smem += 1
        top = v
    assert all(s is None for s in stack)
    return smem


def test_schedule_invariants_on_random_trees_and_dirty_sets():
    rng = random.Random(1)
    nodes = touches = 0
    for it in range(4000):
        n = rng.choice([2, 3, 5, 8, 16, 24, 40, 64])
        father, left, right, root = random_tree(n, rng)
        need = [0] * (2 * n - 1)
        full = it % 2 == 0
        if full:                                         # useOld = 0: every internal node
            for v in range(n, 2 * n - 1):
                need[v] = 1
        else:                                            # a proposal: a few dirty nodes and their ancestors
            for _ in range(rng.randrange(1, 4)):
                v = rng.randrange(2 * n - 1)
                u = father[v] if v < n else v
                while u >= 0 and not need[u]:
                    need[u] = 1
                    u = father[u]
        sched, k, max_depth = build_schedule(n, father, left, right, root, need)
        assert sorted(sched) == list(range(k)) and k == sum(need)
        assert max_depth <= max(1, int(math.log2(n)))    # heavier child first bounds the parked results
        s = walk_schedule(n, left, right, need, sched, k)
        if full and n == 24:
            nodes += k
            touches += s
    # the figure DESIGN.md quotes: about 0.6 shared-memory vector accesses per node in a full evaluation at n = 24
    assert 0.4 < touches / nodes < 0.8


def test_caterpillar_and_balanced_extremes():
    for n in (16, 32):
        # caterpillar: every internal node has a leaf child -> nothing is ever parked
        N = 2 * n - 1
        father, left, right = [-1] * N, [-1] * N, [-1] * N
        prev = 0
        for i, v in enumerate(range(n, N)):
            left[v], right[v] = prev, i + 1
            father[prev] = father[i + 1] = v
            prev = v
        need = [0] * n + [1] * (n - 1)
        sched, k, max_depth = build_schedule(n, father, left, right, prev, need)
        assert max_depth == 0 and walk_schedule(n, left, right, need, sched, k) == 0
        # perfectly balanced: the deepest stack, log2(n) - 1 parked results
        father, left, right = [-1] * N, [-1] * N, [-1] * N
        level, nxt = list(range(n)), n
        while len(level) > 1:
            up = []
            for a, b in zip(level[::2], level[1::2]):
                left[nxt], right[nxt], father[a], father[b] = a, b, nxt, nxt
                up.append(nxt)
                nxt += 1
            level = up
        sched, k, max_depth = build_schedule(n, father, left, right, level[0], need)
        assert max_depth == int(math.log2(n)) - 1
        walk_schedule(n, left, right, need, sched, k)
