#!/usr/bin/env python3
"""small end-to-end runs of every code path for `compute-sanitizer --tool memcheck python tools/sanitize_run.py`
(the pytest suite is too slow under the sanitizer): single volume, 3 z-slabs, both back-ends, atlas, isolevel, raw ingest, post-smooth"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402

eng = lib.Engine(0)
grp = lib.LocalSlabGroup(3)
vols = {"blobs": (synth.random_blobs((30, 37, 41), seed=3), 0.2), "w33": (synth.random_blobs((17, 20, 33), seed=7), 0.1),
        "gyroid": (synth.gyroid(96, P=32), 0.0)}
for name, (vol, iso) in vols.items():
    for backend, omc, ps, ol, fb in ((0, 0, 1, 1, 1), (0, 1, 0, 0, 0), (1, 0, 1, 1, 0), (1, 0, 0, 0, 1)):
        v, t, r = eng.meshify(vol, iso, omc, ps, ol, fb, backend)
        if vol.shape[0] >= 12:
            nz = vol.shape[0]
            gv, gt, _ = grp.meshify(vol, [0, nz // 3, 2 * nz // 3, nz], iso, original_mc=omc, pre_smooth=ps, only_largest=ol,
                                    fill_bubbles=fb, backend=backend)
            assert np.array_equal(gt, t) and np.array_equal(gv.view(np.uint64), v.view(np.uint64)), (name, backend)
        print(name, backend, omc, ps, ol, fb, len(v), len(t), flush=True)
lab = np.zeros((24, 26, 40), np.float32)
lab[3:12, 4:15, 5:30] = 1
lab[14:22, 10:24, 8:20] = 2
lab[0:4, 0:6, 30:40] = 3
d = eng.upload(lab)
for info in eng.atlas_scan(d)[1:]:
    if info.nvox:
        v, t, _ = eng.meshify_label(d, info, 0.5, 0, 1, 1, 0)
        print("label", info.label, len(v), len(t), flush=True)
print("isolevel", [eng.isolevel(d, m) for m in (1, 2, 3)], eng.isolevel(vols["blobs"][0], 2), flush=True)
d.free()
raw = (np.clip(vols["blobs"][0], -2, 2) * 1000).astype(np.int16)
v, t, _ = eng.meshify_raw(raw, 0.2, 0.001, 0.0, [[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 1, 0]], 0, 1, 1, 0, 0)
print("raw", len(v), len(t))
v, t, _ = eng.meshify(vols["gyroid"][0], 0.0, 0, 1, 1, 0, 0)
s = eng.laplacian_hc(v, t, 3)
print("post-smooth", len(v), float(np.abs(s - v).max()), flush=True)
eng.laplacian_hc(np.random.default_rng(0).normal(size=(50, 3)), np.random.default_rng(1).integers(0, 50, size=(120, 3)).astype(np.int32), 2)
grp.close()
print("SANITIZE_RUN_DONE")
