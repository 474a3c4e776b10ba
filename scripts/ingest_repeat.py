import importlib, json, sys, tempfile, os, time
sys.path.insert(0, ".")
gp = importlib.import_module("g-phocs_b200"); synth = importlib.import_module("g-phocs_b200.synth")
for cfg, L in (("hap16", 10000), ("dip8mig", 10000), ("pop6mig4", 20000)):
    m = synth.config(cfg); tmp = tempfile.mkdtemp(); p = os.path.join(tmp, "s.txt")
    synth.generate(m, L, seed=4242, seqfile=p)
    names = synth.sample_slots(m)
    for rep in range(4):
        t0 = time.perf_counter(); a = gp.Alignment.read(p, names); dt = time.perf_counter() - t0
        t = a.timings(); print(cfg, rep, "wall %.1f ms" % (dt * 1e3), {k: round(v * 1e3, 2) for k, v in t.items() if k.endswith("_s")}, a.P, a.U); a.close()
