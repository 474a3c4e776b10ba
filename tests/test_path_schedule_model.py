"""CPU: a Python model of the schedule a locus' team builds for a node move inside k_sweep
(g-phocs_b200/csrc/sweep_kernels.cuh: sweepPathSchedule) and of the column walk that executes it.

A coalescence-time move dirties the moved node and its ancestors; an SPR (executeGenSPR, LocusDataLikelihood.c:931-1012)
the moved father, its old father and their ancestors: one path to the root or two paths that join.  The model restates
the rule — path 1 from the moved node, path 2 from the second node up to (excluding) the first node already on path 1,
longer tail first, other tail, then the stem; a child is a LEAF, a clean vector in HBM (GLOBAL), the vector of the entry
right before (TOP: register top) or the first tail's top parked in the single stack row (STACK) — and then walks the
schedule like a column thread, checking what the kernel relies on:
  * every dirty node is scheduled exactly once, children before parents, the root last,
  * TOP always refers to the entry right before, STACK to the one parked entry, which is still intact when it is read,
  * the values computed equal a plain recursive evaluation of the edited tree (any child order: the node function is
    symmetric, like the product of the two children's factors in computeSubtreeConditionals_new, .c:1650-1673).
The CUDA code is checked bit for bit against the stepwise route on the GPU (tests/test_gpu_sampler.py); this file
documents the algorithm in executable form and guards its corner cases (junction at the moved father, second node on
path 1, moves at and next to the root, caterpillar trees)."""
import random

from test_schedule_model import random_tree

LEAF, STACK, GLOBAL, TOP = 0, 1, 2, 3


def spr(father, left, right, root, sub, target):
    """executeGenSPR on plain arrays; returns (new root, moved father, old grandfather or -1)."""
    F = father[sub]
    G = father[F]
    S = left[F] + right[F] - sub
    TF = father[target]
    if target in (S, F):
        return root, F, G
    father[S] = G
    if G >= 0:
        if left[G] == F:
            left[G] = S
        else:
            right[G] = S
    father[F] = TF
    left[F], right[F] = sub, target
    father[target] = F
    if TF < 0:
        return F, F, G
    if left[TF] == target:
        left[TF] = F
    else:
        right[TF] = F
    return (S if G < 0 else root), F, G


def path_schedule(n, father, left, right, first, second):
    """sweepPathSchedule: list of (node, kindLeft, kindRight, push) in execution order."""
    on_path = set()
    tmp_a, u = [], first
    while u >= 0:
        tmp_a.append(u); on_path.add(u); u = father[u]
    tmp_b, u = [], second
    while u >= n and u not in on_path:
        tmp_b.append(u); u = father[u]
    len_a = tmp_a.index(father[tmp_b[-1]]) if tmp_b else 0
    len_b = len(tmp_b)
    a_first = len_a >= len_b
    order = (tmp_a[:len_a] + tmp_b if a_first else tmp_b + tmp_a[:len_a]) + tmp_a[len_a:]
    len_first = len_a if a_first else len_b
    both = len_a > 0 and len_b > 0
    dirty = set(order)
    sched = []
    for i, v in enumerate(order):
        prev = order[i - 1] if i > 0 else -1
        kinds = []
        for x in (left[v], right[v]):
            kinds.append(LEAF if x < n else GLOBAL if x not in dirty else TOP if x == prev else STACK)
        sched.append((v, kinds[0], kinds[1], both and i == len_first - 1))
    return sched


def node_value(a, b, v):
    return (a * b + 31 * v + 7) % 1000003          # symmetric in its children, like the product of the two factors


def evaluate(n, left, right, x, memo):
    if x < n:
        return x + 1
    if x not in memo:
        memo[x] = node_value(evaluate(n, left, right, left[x], memo), evaluate(n, left, right, right[x], memo), x)
    return memo[x]


def walk(n, left, right, sched, stored):
    """The column thread: register top, one stack row, HBM records (`stored`, updated in place)."""
    top, stack, stack_owner = None, None, None
    done = []
    for i, (v, kl, kr, push) in enumerate(sched):
        vals = []
        for kind, x in ((kl, left[v]), (kr, right[v])):
            if kind == LEAF:
                vals.append(x + 1)
            elif kind == GLOBAL:
                assert x not in done, "a clean child must not have been recomputed in this walk"
                vals.append(stored[x])
            elif kind == TOP:
                assert i > 0 and sched[i - 1][0] == x
                vals.append(top)
            else:
                assert stack_owner == x, "the parked vector is the first tail's top"
                vals.append(stack)
        top = node_value(vals[0], vals[1], v)
        stored[v] = top
        done.append(v)
        if push:
            assert stack is None, "one parking row, used once"
            stack, stack_owner = top, v
    return top


def check_case(n, father, left, right, root, first, second, stored):
    N = 2 * n - 1
    sched = path_schedule(n, father, left, right, first, second)
    order = [e[0] for e in sched]
    # every dirty node once: the two paths to the root
    dirty = set()
    for s in (first, second):
        u = s
        while u >= n:
            dirty.add(u); u = father[u]
    assert sorted(order) == sorted(dirty) and len(order) == len(set(order))
    assert order[-1] == root
    pos = {v: i for i, v in enumerate(order)}
    for v in order:
        for x in (left[v], right[v]):
            if x in pos:
                assert pos[x] < pos[v], "children first"
    assert sum(1 for e in sched if e[3]) <= 1
    got = walk(n, left, right, sched, stored)
    memo = {}
    want = evaluate(n, left, right, root, memo)
    assert got == want
    for v in range(n, N):
        assert stored[v] == memo[v], "every stored vector is current after the walk"


def test_age_moves_schedule_one_path():
    rng = random.Random(5)
    for n in (2, 3, 5, 16, 32):
        for _ in range(40):
            father, left, right, root = random_tree(n, rng)
            memo = {}
            evaluate(n, left, right, root, memo)
            stored = dict(memo)
            node = rng.randrange(n, 2 * n - 1)
            stored_before = dict(stored)
            check_case(n, father, left, right, root, node, -1, stored)
            assert stored == stored_before       # the tree did not change: recomputing the path reproduces what was stored


def test_spr_moves_schedule_two_joining_paths():
    rng = random.Random(11)
    shapes = {"first_is_junction": 0, "second_on_path": 0, "two_tails": 0, "root_change": 0}
    for n in (3, 4, 6, 16, 32):
        for _ in range(300):
            father, left, right, root = random_tree(n, rng)
            memo = {}
            evaluate(n, left, right, root, memo)
            stored = dict(memo)
            N = 2 * n - 1
            sub = rng.choice([x for x in range(N) if x != root])
            F = father[sub]
            # target: any branch outside the pruned subtree and not the father itself
            below = set()
            stack = [sub]
            while stack:
                x = stack.pop(); below.add(x)
                if x >= n:
                    stack += [left[x], right[x]]
            cands = [x for x in range(N) if x not in below and x != F]
            target = rng.choice(cands)
            new_root, F2, G = spr(father, left, right, root, sub, target)
            assert F2 == F
            sched = path_schedule(n, father, left, right, F, G)
            tails = [e for e in sched if e[3]]
            if tails:
                shapes["two_tails"] += 1
            if new_root != root:
                shapes["root_change"] += 1
            u, on1 = F, set()
            while u >= 0:
                on1.add(u); u = father[u]
            if G >= 0 and G in on1:
                shapes["second_on_path"] += 1
            elif G >= 0:
                u = G
                while u not in on1:
                    u = father[u]
                if u == F:
                    shapes["first_is_junction"] += 1
            check_case(n, father, left, right, new_root, F, G, stored)
    assert all(v > 0 for v in shapes.values()), shapes


def test_caterpillar_tree_uses_the_longest_paths():
    n = 32
    N = 2 * n - 1
    father, left, right = [-1] * N, [-1] * N, [-1] * N
    prev = 0
    for v in range(n, N):
        leaf = v - n + 1
        left[v], right[v] = prev, leaf
        father[prev], father[leaf] = v, v
        prev = v
    root = N - 1
    memo = {}
    evaluate(n, left, right, root, memo)
    stored = dict(memo)
    check_case(n, father, left, right, root, n, -1, stored)            # the deepest node: a path of n - 1 entries
    assert len(path_schedule(n, father, left, right, n, -1)) == n - 1
