#!/usr/bin/env python3
"""tests/golden/golden_post.json: sha256 of the vertex array after the UNMODIFIED reference's laplacian_smoothHC()
(src/quadric.c:343-394, compiled into oracle/_ref by oracle/build_ref.sh) on meshes the reference's meshify() makes
from the named parity volumes (tests/cases.py).  Run here, where /root/reference exists:  python tools/make_golden_post.py"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402

CASES = [("sphere40", (0, 1, 1, 0)), ("blobs", (0, 0, 0, 0)), ("blobs2", (0, 1, 1, 1)), ("gyroid96", (0, 1, 0, 1)), ("bet", (0, 1, 1, 0))]
RUNS = [(1, True), (3, True), (3, False), (10, True)]


def main():
    R = oracle.Ref("lewiner")
    vols = cases.volumes()
    out = {}
    for name, (omc, p, l, b) in CASES:
        vol, iso = vols[name]
        m = R.meshify(vol, iso, omc, p, l, b)
        assert m["rc"] == 0
        v, t = m["verts"], m["tris"]
        key = f"{name}/o{omc}p{p}l{l}b{b}"
        out[key] = {"nverts": len(v), "ntris": len(t), "mesh": hashlib.sha256(v.tobytes() + t.tobytes()).hexdigest()}
        for it, lock in RUNS:
            s = R.laplacian_hc(v, t, it, lock_edges=lock)
            out[key][f"iter{it}_lock{int(lock)}"] = hashlib.sha256(np.ascontiguousarray(s).tobytes()).hexdigest()
        print(key, len(v), len(t))
    (ROOT / "tests" / "golden" / "golden_post.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
