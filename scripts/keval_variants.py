#!/usr/bin/env python
"""Times the full data-likelihood launch (k_eval) of one library build on the bench workload.

    python scripts/keval_variants.py LIB.so [config] [loci]

Development tool: `LIB.so` is a build of csrc/ with other compile-time settings (GPHOCS_KSTACK,
GPHOCS_EVAL_MINBLOCKS); GPHOCS_EVAL_SMEM_BUDGET in the environment caps the loci per CTA batch.  Prints one line.
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
gp = importlib.import_module("g-phocs_b200")
gp.LIB_PATH = os.path.abspath(sys.argv[1])
synth = importlib.import_module("g-phocs_b200.synth")
cfg = sys.argv[2] if len(sys.argv) > 2 else "pop6mig4"
L = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
w = synth.generate(synth.config(cfg), L, seed=1000)
stream = torch.cuda.Stream()
st = gp.LociStore.from_workload(w, device=0, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    for _ in range(5):
        st.evaluate_device(0)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(100):
        st.evaluate_device(0)
    b.record(stream)
torch.cuda.synchronize()
full = a.elapsed_time(b) / 100
print(f"{os.path.basename(sys.argv[1])} budget={os.environ.get('GPHOCS_EVAL_SMEM_BUDGET', '-')} {cfg} {L}: full {1e3 * full:.1f} us "
      f"sum {float(st.evaluate(0).sum()):.10f}")
st.close()
