#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  python scripts/launch_summary.py FILE.csv"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
tot = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    v = float(r[-1].replace(",", ""))
    v = v / 1000 if r[-2] == "ns" else v * 1000 if r[-2] == "ms" else v
    t = tot.setdefault(name, [0.0, 0])
    t[0] += v; t[1] += 1
s = sum(t[0] for t in tot.values())
for k, (t, c) in sorted(tot.items(), key=lambda x: -x[1][0]):
    print(f"{k:34s} n={c:4d}  total {t:10.1f} us  avg {t / c:8.1f} us  {100 * t / s:5.1f}%")
print(f"{len(rows)} launches, {s:.1f} us")
