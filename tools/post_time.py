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
for rep in range(2):
    eng.set_profile(rep == 1)
    eng.sync()
    t0 = time.perf_counter()
    eng.laplacian_hc_result(r, iters)
    ms = (time.perf_counter() - t0) * 1e3
print(f"G{n}: {r.nverts} verts {r.ntris} tris, {iters} iterations: {ms:.2f} ms on the device")
per = {}
for k, x in eng.kernel_times():
    per[k] = per.get(k, 0.0) + x
print({k: round(x, 3) for k, x in per.items()})
if v is not None:
    import oracle
    if oracle.ref_available("lewiner"):
        R = oracle.Ref("lewiner")
        t0 = time.perf_counter()
        R.laplacian_hc(v, t, iters)
        print(f"reference laplacian_smoothHC on the same mesh, 1 core: {(time.perf_counter() - t0) * 1e3:.0f} ms")
