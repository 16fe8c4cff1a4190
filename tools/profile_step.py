#!/usr/bin/env python3
"""two whole meshify() steps on G<size> (config 3 flags) for ncu: the second step is the warm one.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:'k_smooth3|k_mc_emit|...' -o gpurun_out/full python tools/profile_step.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = lib.Engine(0)
d = eng.tiled_volume(synth.gyroid_tile(128), (n, n, n))
for _ in range(steps):
    _, _, r = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
print(r.nverts, r.ntris, r.times())
