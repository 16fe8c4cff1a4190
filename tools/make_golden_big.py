#!/usr/bin/env python3
"""Generate tests/golden/golden_big.json from the UNMODIFIED reference (oracle/_ref/libref_lewiner.so), in the build
container where /root/reference exists (minutes of CPU time):

    python tools/make_golden_big.py [--threads 8]

  gyroid     G256 / G512 (BASELINE config 3 family, Lewiner -p1 -l1 -b1 iso 0): counts + topology digests of the
             reference's meshes, so that the full-size GPU runs are pinned by topology, not only by counts
  boxlaw     the reference's PRE-weld counts on non-cubic G volumes of (a, b, c) 128-voxel tiles (x, y, z): an exact
             fit  nv = A abc + B1 ab + B2 bc + B3 ca + C1 a + C2 b + C3 c + D  on the small boxes, checked on held-out
             larger ones -> the known answer of the multi-GPU bench volumes (2048x1024x1024, 2048x2048x1024) that no
             CPU run can reach
  atlas      D99 (BASELINE config 4): for EVERY label 1..nlabel the voxel count and, for the non-empty ones, the
             reference's per-label mesh counts, digest and whether the isolevel was reset (src/nii2mesh.c:492-583
             restated around the reference's own meshify(): whole-volume binarisation, iso 0.5, -l off, -p1 -b0)
"""
import argparse
import json
import sys
import multiprocessing as mp
from fractions import Fraction
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import synth  # noqa: E402
from oracle import Ref, build  # noqa: E402
from oracle.canon import topology_digest  # noqa: E402

OUT = ROOT / "tests" / "golden" / "golden_big.json"


def box_volume(a, b, c):
    """G volume of a x b x c tiles along x, y, z -> array [z, y, x]"""
    return np.tile(synth.gyroid_tile(128), (c, b, a))


def solve_exact(rows, rhs):
    """exact rational least-squares is overkill: solve the square system of the first len(rows[0]) rows exactly"""
    n = len(rows[0])
    M = [[Fraction(x) for x in r] + [Fraction(y)] for r, y in zip(rows[:n], rhs[:n])]
    for i in range(n):
        p = next(r for r in range(i, n) if M[r][i] != 0)
        M[i], M[p] = M[p], M[i]
        M[i] = [x / M[i][i] for x in M[i]]
        for r in range(n):
            if r != i and M[r][i] != 0:
                M[r] = [x - M[r][i] * y for x, y in zip(M[r], M[i])]
    return [M[i][n] for i in range(n)]


def terms(a, b, c):
    return [a * b * c, a * b, b * c, c * a, a, b, c, 1]


def _box_job(box):
    o = Ref("lewiner").meshify(box_volume(*box), 0.0, 0, 1, 1, 1, counts=True)
    r = (o["pre_nv"], o["pre_nt"], len(o["verts"]), len(o["tris"]))
    sys.stderr.write(f"box {box} {r}\n")
    return r


_D99 = None


def _label_job(lab):
    global _D99
    if _D99 is None:
        _D99 = synth.load_nifti(ROOT / "tests" / "golden" / "D99_atlas_v2.0_right.nii.gz")[0]
    vol = _D99
    b = ((vol > np.float32(lab - 0.5)) & (vol < np.float32(lab + 0.5))).astype(np.float32)
    n1 = int(b.sum())
    if n1 == 0:
        return dict(nvox=0)
    o = Ref("lewiner").meshify(b, 0.5, 0, 1, 0, 0, counts=True)
    assert o["rc"] == 0, lab
    nu, nt, dg = topology_digest(o["verts"], o["tris"])
    if lab % 20 == 0:
        sys.stderr.write(f"atlas label {lab}\n")
    return dict(nvox=n1, nverts=len(o["verts"]), ntris=len(o["tris"]), nused=nu, digest=dg, iso_reset=int(o["iso_reset"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--skip-atlas", action="store_true")
    args = ap.parse_args()
    build()
    out = json.loads(OUT.read_text()) if OUT.exists() else {}

    # ---- gyroid digests ----
    R = Ref("lewiner")
    out.setdefault("gyroid", {})
    for n in (256, 512):
        if str(n) in out["gyroid"]:
            continue
        r = R.meshify(np.tile(synth.gyroid_tile(128), (n // 128,) * 3), 0.0, 0, 1, 1, 1)
        assert r["rc"] == 0
        nu, nt, dg = topology_digest(r["verts"], r["tris"])
        _, _, dgt = topology_digest(r["verts"], r["tris"], with_coords=False)
        out["gyroid"][str(n)] = dict(nverts=len(r["verts"]), ntris=len(r["tris"]), nused=nu, digest=dg, faces_digest=dgt)
        print("gyroid", n, out["gyroid"][str(n)], flush=True)
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")

    # ---- box law of the pre-weld counts ----
    if "boxlaw" not in out:
        fit_boxes = [(1, 1, 1), (2, 1, 1), (1, 2, 1), (1, 1, 2), (2, 2, 1), (1, 2, 2), (2, 1, 2), (2, 2, 2), (3, 1, 2), (1, 3, 2)]
        hold = [(3, 2, 1), (2, 3, 3), (4, 2, 2), (4, 4, 2), (3, 3, 3), (4, 3, 5)]
        with mp.Pool(min(args.threads, 4)) as pool:
            samples = dict(zip(fit_boxes + hold, pool.map(_box_job, fit_boxes + hold, chunksize=1)))
        law = {}
        for k, name in ((0, "pre_nverts"), (1, "pre_ntris")):
            coef = solve_exact([terms(*b) for b in fit_boxes], [samples[b][k] for b in fit_boxes])
            for b in fit_boxes + hold:
                pred = sum(cf * t for cf, t in zip(coef, terms(*b)))
                assert pred == samples[b][k], (name, b, pred, samples[b][k])
            assert all(cf.denominator == 1 for cf in coef), coef
            law[name] = [int(cf) for cf in coef]
        out["boxlaw"] = dict(terms="abc ab bc ca a b c 1  (a, b, c = 128-voxel tiles along x, y, z)", **law,
                             samples={"x".join(map(str, b)): list(v) for b, v in samples.items()},
                             checked_on=["x".join(map(str, b)) for b in hold])
        print("boxlaw", law, flush=True)
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")

    # ---- D99: every label ----
    if not args.skip_atlas and "atlas" not in out:
        vol, _ = synth.load_nifti(ROOT / "tests" / "golden" / "D99_atlas_v2.0_right.nii.gz")
        nlabel = int(np.trunc(vol.max()))
        with mp.Pool(args.threads) as pool:
            res = dict(zip(range(1, nlabel + 1), pool.map(_label_job, range(1, nlabel + 1), chunksize=4)))
        out["atlas"] = dict(nlabel=nlabel, flags="p1 l0 b0 iso 0.5", nonempty=sum(1 for v in res.values() if v["nvox"]),
                            resets=sorted(k for k, v in res.items() if v.get("iso_reset")),
                            labels={str(k): res[k] for k in sorted(res)})
        print("atlas: nonempty", out["atlas"]["nonempty"], "resets", out["atlas"]["resets"], flush=True)
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
