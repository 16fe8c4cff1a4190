#!/usr/bin/env python3
"""post-smooth (-s) timing: G<size> mesh on the device, b2m_laplacian_hc_device in place, per-kernel times;
the reference's laplacian_smoothHC on the same mesh when oracle/_ref is present and the mesh is small enough
    python tools/post_time.py [size] [iters]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng = lib.Engine(0)
d = eng.tiled_volume(synth.gyroid_tile(128), (n, n, n))
_, _, r = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
v, t = eng.fetch(r) if n <= 512 else (None, None)
eng.laplacian_hc_result(r, iters)          # warm-up run; its result is the one compared with the reference
g = eng.fetch(r)[0] if n <= 512 else None
_, _, r = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)   # fresh mesh for the timed run
eng.set_profile(True)
eng.sync()
t0 = time.perf_counter()
eng.laplacian_hc_result(r, iters)
ms = (time.perf_counter() - t0) * 1e3
import json
per = {}
for k, x in eng.kernel_times():
    per[k] = per.get(k, 0.0) + x
out = {"workload": f"G{n} gyroid+bumps -p1 -l1 -b1 mesh, laplacian_smoothHC alpha 0.1 beta 0.5 lockEdges, {iters} iterations (nii2mesh -s {iters})",
       "nverts": r.nverts, "ntris": r.ntris, "device_ms": round(ms, 2), "kernels_ms": {k: round(x, 3) for k, x in per.items()},
       "reference_ms_1core": None}
if v is not None:
    import oracle
    if oracle.ref_available("lewiner"):
        R = oracle.Ref("lewiner")
        t0 = time.perf_counter()
        ref = R.laplacian_hc(v, t, iters)
        out["reference_ms_1core"] = round((time.perf_counter() - t0) * 1e3)
        out["bit_identical_to_reference"] = bool((g.view("u8") == ref.view("u8")).all())
print(json.dumps(out))
