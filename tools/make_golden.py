#!/usr/bin/env python3
"""Generate tests/golden/golden.json from the UNMODIFIED reference compiled by oracle/build_ref.sh
(oracle/_ref/libref_{lewiner,classic}.so, mctest).  Run in the build container, where
/root/reference exists:

    python tools/make_golden.py

Every entry is an order-independent digest (oracle/canon.py:topology_digest) or a byte checksum of
what the reference itself produced for a deterministic input that the tests can regenerate
(nii2mesh_b200/synth.py, tests/surfaces.py, tests/golden/bet.nii.gz).  The tests then hold the
oracle restatement (CPU) and the CUDA path (GPU) to these values without needing the reference.
"""
import hashlib
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from nii2mesh_b200 import synth  # noqa: E402
from oracle import Ref, build, REF_DIR  # noqa: E402
from oracle.canon import topology_digest  # noqa: E402
import surfaces  # noqa: E402
import cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    build()
    RL, RC = Ref("lewiner"), Ref("classic")
    out = {"meshify": {}, "smooth": {}, "bwlabel": {}, "dilate": {}, "surfaces": {}, "front": {}}

    # 1. the self-test surfaces: our numpy restatement must equal the volumes mctest writes
    with tempfile.TemporaryDirectory() as td:
        log = subprocess.run([str(REF_DIR / "mctest")], cwd=td, capture_output=True, text=True).stdout
        counts = [tuple(int(t) for t in (ln.split()[3], ln.split()[5])) for ln in log.splitlines()
                  if ln.startswith("output mesh vert")]
        assert counts == surfaces.KNOWN, counts
        for k in range(10):
            raw = np.fromfile(Path(td) / f"{k}.nii", dtype=np.float32, offset=352).reshape(60, 60, 60)
            mine = surfaces.surface(k)
            nd = int((raw.view(np.uint32) != mine.view(np.uint32)).sum())
            assert nd == 0, f"surface {k}: numpy restatement differs from mctest in {nd} voxels"
            v, t = RL.mc(mine, [0, 0, 0], [59, 59, 59], 0.0, 0)
            assert (len(v), len(t)) == surfaces.KNOWN[k], (k, len(v), len(t))
            out["surfaces"][str(k)] = dict(nv=len(v), nt=len(t), digest=topology_digest(v, t)[2])
        v, t = RL.mc(surfaces.surface(7), [0, 0, 0], [59, 59, 59], 0.0, 1)
        assert (len(v), len(t)) == surfaces.KNOWN_ORIGINAL[7]
        out["surfaces"]["7_original"] = dict(nv=len(v), nt=len(t), digest=topology_digest(v, t)[2])

    # 2. stage checksums and whole-path digests on the named volumes
    for name, (vol, iso) in cases.volumes().items():
        out["smooth"][name] = sha(RL.smooth(vol))
        mask = (vol >= np.float32(iso)).astype(np.float32)
        for ol, fb in ((1, 0), (0, 1), (1, 1)):
            out["bwlabel"][f"{name}/l{ol}b{fb}"] = sha(RL.bwlabel(mask, 18, ol, fb) != 0)
        out["dilate"][name] = sha(RL.dilate(mask) != 0)
        for backend, omc, ps, ol, fb in cases.flag_sets(name):
            R = RC if backend == 1 else RL
            r = R.meshify(vol, iso, omc, ps, ol, fb, return_img=True)
            key = f"{name}/backend{backend}_o{omc}_p{ps}_l{ol}_b{fb}"
            if r["rc"] != 0:
                out["meshify"][key] = dict(rc=r["rc"])
                continue
            nu, nt, dg = topology_digest(r["verts"], r["tris"])
            _, _, dgt = topology_digest(r["verts"], r["tris"], with_coords=False)
            out["meshify"][key] = dict(rc=0, nverts=len(r["verts"]), nused=nu, ntris=nt, digest=dg, faces_digest=dgt)
            out["front"][key] = sha(r["img"])  # the reference's mutated img = the composed volume
            print(key, out["meshify"][key]["nverts"], nt, flush=True)
    (ROOT / "tests" / "golden" / "golden.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote tests/golden/golden.json")


if __name__ == "__main__":
    main()
