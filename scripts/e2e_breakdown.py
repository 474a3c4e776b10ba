"""Times the pieces of one end-to-end step (host buffers -> C ABI -> host results) on the GPU box."""
import importlib, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("g-phocs_b200")
synth = importlib.import_module("g-phocs_b200.synth")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = synth.generate(synth.config("pop6mig4"), L, seed=1000)
st = gp.LociStore.from_workload(w)
gen = gp.Genealogy(L, w.pops)
hw = {k: gp.pinned_like(getattr(w, k)) for k in ("father", "left", "right", "age", "root", "ev_start", "pop_start", "ev_type", "ev_id", "ev_time")}
out = gp.pinned_like(np.zeros(L))
def t(f, n=5):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("set_trees pinned  ms", t(lambda: st.set_trees(hw["father"], hw["left"], hw["right"], hw["age"], hw["root"])))
print("set_trees pageable ms", t(lambda: st.set_trees(w.father, w.left, w.right, w.age, w.root)))
print("evaluate(0)+D2H   ms", t(lambda: st.evaluate(0, want_sum=True, out=out)))
print("set_events pinned ms", t(lambda: gen.set_events(hw["ev_start"], hw["pop_start"], hw["ev_type"], hw["ev_id"], hw["ev_time"])))
print("gen.evaluate      ms", t(lambda: gen.evaluate(per_locus_stats=False)))
ops = gp.make_ops(np.arange(L), gp.OP_ADJUST_AGE, a=w.n + 2, x=w.age[:, w.n + 2] * 1.001)
rej = gp.make_ops(np.arange(L), gp.OP_REVERT)
st.apply_ops(gp.make_ops(np.arange(L), gp.OP_COMMIT))
print("apply_ops(adjust) ms", t(lambda: (st.apply_ops(ops), st.apply_ops(rej))[0]) / 2)
def cyc():
    st.apply_ops(ops); st.evaluate(1, out=out); st.apply_ops(rej)
print("cycle             ms", t(cyc))
st.apply_ops(ops)
print("evaluate(1)+D2H   ms", t(lambda: st.evaluate(1, out=out)))
