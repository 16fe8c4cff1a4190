#!/usr/bin/env python3
"""dev diagnostic: classic back-end positions, GPU vs oracle (bitwise)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nii2mesh_b200 import lib
from oracle import Oracle
import cases
E = lib.Engine(0); O = Oracle()
for name in ("gyroid96", "sphere40", "blobs"):
    vol, iso = cases.volumes(big=False)[name]
    gv, gt, r = E.meshify(vol, iso, 0, 0, 0, 0, 1)
    o = O.meshify(vol, iso, 0, 0, 0, 0, 1)
    ov = o["verts"]
    from scipy.spatial import cKDTree
    d, j = cKDTree(ov).query(gv)
    diff = np.abs(gv - ov[j])
    bad = np.nonzero((gv != ov[j]).any(axis=1))[0]
    print(name, len(gv), len(ov), "non-bit-identical:", len(bad), "max diff", diff.max())
    for b in bad[:8]:
        print("   gpu", gv[b].tolist(), "ref", ov[j[b]].tolist(), "d", (gv[b] - ov[j[b]]).tolist())
    ax = (gv[bad] != ov[j[bad]])
    print("   per-axis mismatch counts", ax.sum(axis=0).tolist())
