#!/usr/bin/env python3
"""Generate tests/golden/atlas_golden.json from the UNMODIFIED reference (oracle/_ref/libref_lewiner.so): the
reference's atlas loop (/root/reference/src/nii2mesh.c:492-583) restated around its own meshify() - whole-volume
binarisation per label, isolevel 0.5, -l off - on data/D99_atlas_v2.0_right.nii.gz (copied to tests/golden/) for a
handful of labels (each costs the reference ~1-2 s: it smooths all 23 M voxels per label).

    python tools/make_golden_atlas.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import synth  # noqa: E402
from oracle import Ref, build  # noqa: E402
from oracle.canon import topology_digest  # noqa: E402

LABELS = [1, 2, 7, 33, 100, 163, 164, 250, 301, 365]   # incl. labels without voxels (skipped by the reference)


def main():
    build()
    R = Ref("lewiner")
    vol, _ = synth.load_nifti(ROOT / "tests" / "golden" / "D99_atlas_v2.0_right.nii.gz")
    nlabel = int(np.trunc(vol.max()))
    out = {"nlabel": nlabel, "shape": list(vol.shape), "labels": {}}
    counts = np.bincount(np.rint(vol).astype(np.int64).ravel(), minlength=nlabel + 1)
    out["nonempty"] = int((counts[1:] > 0).sum())
    for lab in LABELS:
        b = ((vol > np.float32(lab - 0.5)) & (vol < np.float32(lab + 0.5))).astype(np.float32)
        n1 = int(b.sum())
        if n1 == 0:
            out["labels"][str(lab)] = {"nvox": 0}
            continue
        for ps, fb in ((1, 0), (1, 1)):
            o = R.meshify(b, 0.5, 0, ps, 0, fb)
            assert o["rc"] == 0
            nu, nt, dig = topology_digest(o["verts"], o["tris"])
            out["labels"].setdefault(str(lab), {"nvox": n1})[f"p{ps}_b{fb}"] = dict(
                nverts=len(o["verts"]), ntris=len(o["tris"]), nused=nu, digest=dig)
            print(lab, n1, ps, fb, len(o["verts"]), len(o["tris"]), flush=True)
    (ROOT / "tests" / "golden" / "atlas_golden.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
