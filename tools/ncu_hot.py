#!/usr/bin/env python3
"""hot-loop summary of one kernel from an .ncu-rep (source page): opcode mix per loop trip, stall reasons, top stalls
    python tools/ncu_hot.py gpurun_out/x.ncu-rep k_smooth3 [--sass]"""
import csv
import subprocess
import sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2]
first = data[0][0]
# the page lists the function once per profiled launch: keep the first copy
for i in range(1, len(data)):
    if data[i][0] == first:
        data = data[:i]
        break
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[isamp]) for r in data)
mx = max(int(r[ia]) for r in data)
print(f"instructions {tot}  samples {ts}  max per-instruction count {mx}  -> {tot / mx:.1f} instr per trip")
c, s = Counter(), Counter()
for r in data:
    op = [o for o in r[1].split() if not o.startswith("@")][0].split(".")[0].rstrip(";")
    c[op] += int(r[ia]); s[op] += int(r[isamp])
for op, n in c.most_common(24):
    print(f"  {op:10s} {n / mx:7.2f}/trip {n / tot * 100:6.2f}%  samples {s[op] / ts * 100:6.2f}%")
for name in hdr:
    if name.startswith("stall_") and "Not" not in name:
        i = hdr.index(name); v = sum(int(r[i]) for r in data)
        if v / ts > 0.01:
            print(f"  {name:24s} {v / ts * 100:5.1f}%")
print("top stall sites:")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:12]:
    print(f"  {int(r[isamp]) / ts * 100:5.1f}%  {r[1].strip()}")
if "--sass" in sys.argv:
    for r in data:
        print(f"{int(r[ia]) / mx:5.2f} {int(r[isamp]):6d}  {r[1].strip()}")
