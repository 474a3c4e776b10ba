#!/usr/bin/env python
"""Instructions / stall samples / active threads of an `ncu --set full --import-source on` capture, aggregated over
source-line regions.   python scripts/ncu_regions.py REPORT.ncu-rep  file:lo-hi=name ...   (unlisted lines: by file)"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr, fname = None, ""
I, S, T = collections.Counter(), collections.Counter(), collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    d = dict(zip(hdr, r))
    try:
        ln = int(d["Line No"]); ins = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0); ti = int(d["Thread Instructions Executed"] or 0)
    except ValueError:
        continue
    key = fname
    for f, lo, hi, name in regions:
        if f == fname and lo <= ln <= hi:
            key = name; break
    I[key] += ins; S[key] += s; T[key] += ti
tot, ts = sum(I.values()), sum(S.values())
for k, v in I.most_common():
    print(f"{k:24s} inst {v/1e6:9.1f}M {100*v/tot:5.1f}%   samples {100*S[k]/max(ts,1):5.1f}%   threads/inst {T[k]/max(v,1):5.1f}")
print(f"total {tot/1e6:.1f}M warp instructions, {ts} samples")
