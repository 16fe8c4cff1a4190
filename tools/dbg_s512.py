#!/usr/bin/env python3
"""dev diagnostic: S512 classic — find the vertices the reference keeps apart and we merge"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nii2mesh_b200 import lib, synth
from oracle import Oracle, Ref, ref_available
from scipy.spatial import cKDTree
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
E = lib.Engine(0)
vol = synth.noisy_sphere(n)
gv, gt, r = E.meshify(vol, 0.0, 1, 0, 0, 0, 1)
print("gpu", len(gv), len(gt), r.pre_nverts, r.pre_ntris, r.nmerged, r.ndegenerate, flush=True)
if ref_available("classic"):
    o = Ref("classic").meshify(vol, 0.0, 1, 0, 0, 0)
else:
    o = Oracle().meshify(vol, 0.0, 1, 0, 0, 0, 1)
ov, ot = o["verts"], o["tris"]
print("ref", len(ov), len(ot), flush=True)
tg = cKDTree(gv)
d, j = tg.query(ov, k=1)
cnt = np.bincount(j, minlength=len(gv))
dup = np.nonzero(cnt > 1)[0]
print("gpu vertices matched by >1 ref vertex:", len(dup), "max nn dist", d.max())
p0 = None
for g in dup[:10]:
    rs = np.nonzero(j == g)[0]
    print(" gpu", g, gv[g].tolist())
    for k in rs:
        print("    ref", k, ov[k].tolist(), "dist", float(np.linalg.norm(ov[k] - gv[g])))
    if len(rs) == 2:
        a, b = ov[rs[0]], ov[rs[1]]
        print("    ref pair distance", float(np.linalg.norm(a - b)))
tr = cKDTree(ov)
d2, j2 = tr.query(gv, k=1)
print("gpu->ref max nn dist", d2.max(), "far:", int((d2 > 1e-5).sum()))
