#!/usr/bin/env python
"""Warp-instruction and stall-sample totals per source-line range of an ncu capture.
    python scripts/ncu_phases.py rep.ncu-rep file.cuh name:lo-hi name:lo-hi ..."""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = {}
for a in sys.argv[3:]:
    k, r = a.split(":"); lo, hi = r.split("-"); ranges[k] = (int(lo), int(hi))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = None; curfile = ""
inst = {}; samp = {}
for r in rows:
    if r and r[0] == "File Path": curfile = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] == "-":
        cur = (curfile.split("/")[-1], int(r[0])); continue
    try:
        n = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    inst[cur] = inst.get(cur, 0) + n; samp[cur] = samp.get(cur, 0) + sm
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp-instr {ti/1e6:.1f}M samples {ts}")
acc_i = acc_s = 0
for k, (lo, hi) in ranges.items():
    i = sum(v for (f, l), v in inst.items() if f == fname and lo <= l <= hi)
    s = sum(v for (f, l), v in samp.items() if f == fname and lo <= l <= hi)
    acc_i += i; acc_s += s
    print(f"{k:12s} instr {i/1e6:7.1f}M ({100*i/ti:4.1f}%)  samples {100*s/max(ts,1):5.1f}%")
print(f"{'other':12s} instr {(ti-acc_i)/1e6:7.1f}M ({100*(ti-acc_i)/ti:4.1f}%)  samples {100*(ts-acc_s)/max(ts,1):5.1f}%")
