#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo).
    python scripts/sass_lines.py g-phocs_b200/csrc/libgphocs_b200.so k_sweep [top N]"""
import collections, os, re, subprocess, sys, tempfile
so, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host_runtime" not in f][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
cur, on, cnt = None, False, collections.Counter()
for line in out.splitlines():
    if line.startswith("//-----") and ".text." in line:
        on = kern in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line) and cur:
        cnt[cur] += 1
tot = sum(cnt.values())
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print(tot, "SASS instructions;", byfile.most_common(10))
for (f, l), c in cnt.most_common(top):
    print(f"{c:5d}  {f}:{l}")
