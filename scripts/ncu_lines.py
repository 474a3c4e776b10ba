#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` capture.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top N]

Reads the `--page source --print-source cuda,sass` CSV and prints, for the lines that collected the most warp-stall
samples, the sample share, instructions executed and the dominant stall reasons.  Needs -lineinfo at compile time.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    fname = ""
    agg = {}
    total = 0
    seen_kernel = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            seen_kernel += 1
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue      # SASS rows carry an address in column 2; source rows carry "-"
        if seen_kernel > 1 and "--all" not in sys.argv:
            pass
        try:
            samples = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        inst = int(r[hdr.index("Instructions Executed")] or 0)
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    v = int(r[i])
                except ValueError:
                    v = 0
                if v:
                    stalls[h[6:]] = v
        key = (fname, int(r[0]))
        a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0, "inst": 0, "stalls": {}})
        a["samples"] += samples
        a["inst"] += inst
        for k, v in stalls.items():
            a["stalls"][k] = a["stalls"].get(k, 0) + v
        total += samples
    print(f"total samples {total}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{k} {v}" for k, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:4])
        print(f"{100.0 * a['samples'] / max(total, 1):5.1f}%  {f}:{ln:<4d} inst {a['inst']:>10d}  [{st}]  {a['src'][:90]}")


if __name__ == "__main__":
    main()
