#!/usr/bin/env python3
"""per-kernel CUDA-event times of the G<size> step (config 3 flags), median over <steps> warm steps:
    python tools/ktime.py [size] [steps] [kernel name filter ...]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
filt = sys.argv[3:]
eng = lib.Engine(0)
d = eng.tiled_volume(synth.gyroid_tile(128), (n, n, n))
for _ in range(2):
    _, _, r = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
tot = []
for _ in range(steps):
    _, _, r = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
    tot.append(r.times()["total"])
eng.set_profile(True)
acc = {}
for _ in range(steps):
    eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
    per = {}
    for k, v in eng.kernel_times():
        per[k] = per.get(k, 0.0) + v
    for k, v in per.items():
        acc.setdefault(k, []).append(v)
eng.set_profile(False)
tot.sort()
print("total_ms median %.3f min %.3f  nv %d nt %d" % (tot[len(tot) // 2], tot[0], r.nverts, r.ntris))
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    if filt and not any(f in k for f in filt):
        continue
    v.sort()
    print("%-18s %.4f" % (k, v[len(v) // 2]))
