"""GPU tests of the atlas front-end (reference: src/nii2mesh.c:492-583) through the C ABI:
b2m_atlas_scan + b2m_meshify_label_device must give, for every label, the mesh the reference gets from the WHOLE binary
volume of that label (isolevel 0.5, -l off) - although only the label's bounding box is ever processed."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.canon import assert_same_mesh, topology_digest

pytestmark = pytest.mark.gpu


def synthetic_atlas():
    """labels 1..9 in a 46 x 50 x 58 volume: blobs at the volume faces and corners, a label in two pieces, a hollow one
    (a bubble for -b), a one-voxel-thick sheet, a label id without voxels (4), non-integer values inside the +-0.5 band
    and exactly on its edge"""
    rng = np.random.default_rng(5)
    nz, ny, nx = 46, 50, 58
    v = np.zeros((nz, ny, nx), np.float32)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")

    def ball(c, r):
        return (z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 <= r * r
    v[ball((20, 25, 30), 9)] = 1
    v[ball((20, 25, 30), 4)] = 0              # cavity inside label 1 -> a bubble
    v[ball((3, 4, 5), 6)] = 2                  # cut by three volume faces
    v[ball((40, 44, 52), 7)] = 3               # cut by the opposite faces
    v[ball((35, 10, 12), 4)] = 5
    v[ball((8, 40, 45), 3)] = 5                # label 5 in two pieces
    v[10:30, 36:37, 8:28] = 6                  # a sheet, one voxel thick
    v[ball((30, 30, 48), 5)] = 7.3             # inside the band of 7
    v[ball((12, 12, 40), 4)] = 8.5             # exactly on the edge: belongs to no label
    v[0:3, :, :][v[0:3, :, :] == 0] = 9        # a slab that covers a whole volume face
    v[ball((22, 8, 50), 3)] = 9
    noise = rng.uniform(-0.2, 0.2, v.shape).astype(np.float32)
    v = np.where(v > 0, v + noise * (v != 8.5), 0).astype(np.float32)
    return v


def test_atlas_scan_counts_and_boxes(eng):
    vol = synthetic_atlas()
    d = eng.upload(vol)
    try:
        infos = eng.atlas_scan(d)
    finally:
        d.free()
    nlabel = int(np.trunc(vol.max()))
    assert len(infos) == nlabel + 1
    for i in range(1, nlabel + 1):
        m = (vol > np.float32(i - 0.5)) & (vol < np.float32(i + 0.5))
        assert infos[i].nvox == int(m.sum()), i
        if m.any():
            zz, yy, xx = np.nonzero(m)
            assert list(infos[i].lo) == [xx.min(), yy.min(), zz.min()], i
            assert list(infos[i].hi) == [xx.max(), yy.max(), zz.max()], i
    assert infos[4].nvox == 0 and infos[8].nvox == 0  # absent id; values exactly on the band edge


@pytest.mark.parametrize("flags", [(0, 0, 1, 0), (0, 0, 1, 1), (0, 0, 0, 0), (0, 1, 1, 0), (1, 0, 1, 1), (1, 0, 0, 0)])
def test_atlas_labels_equal_whole_volume_reference(eng, orc, flags):
    backend, omc, ps, fb = flags
    vol = synthetic_atlas()
    d = eng.upload(vol)
    try:
        infos = eng.atlas_scan(d)
        done = 0
        for info in infos[1:]:
            if info.nvox == 0:
                continue
            b = ((vol > np.float32(info.label - 0.5)) & (vol < np.float32(info.label + 0.5))).astype(np.float32)
            o = orc.meshify(b, 0.5, omc, ps, 0, fb, backend)
            tag = f"label {info.label} flags {flags}"
            if o["rc"] != 0:
                from nii2mesh_b200 import lib
                with pytest.raises(lib.MeshifyFailure):
                    eng.meshify_label(d, info, 0.5, omc, ps, fb, backend)
                continue
            gv, gt, r = eng.meshify_label(d, info, 0.5, omc, ps, fb, backend)
            assert (len(gv), len(gt)) == (len(o["verts"]), len(o["tris"])), tag
            assert (r.pre_nverts, r.pre_ntris) == (o["pre_nv"], o["pre_nt"]), tag
            assert_same_mesh(gv, gt, o["verts"], o["tris"], 1e-5)
            if backend == 0:  # Lewiner: positions bit for bit
                assert topology_digest(gv, gt)[2] == topology_digest(o["verts"], o["tris"])[2], tag
            done += 1
        assert done >= 6
    finally:
        d.free()


def test_atlas_d99_golden(eng):
    """BASELINE configs[3]: data/D99_atlas_v2.0_right.nii.gz; digests recorded from the unmodified reference
    (tools/make_golden_atlas.py) for a handful of labels"""
    from nii2mesh_b200 import synth
    gold = json.loads((GOLDEN / "atlas_golden.json").read_text())
    vol, _ = synth.load_nifti(GOLDEN / "D99_atlas_v2.0_right.nii.gz")
    assert list(vol.shape) == gold["shape"]
    d = eng.upload(vol)
    try:
        infos = eng.atlas_scan(d)
        assert len(infos) == gold["nlabel"] + 1
        assert sum(1 for i in infos[1:] if i.nvox > 0) == gold["nonempty"]
        for lab, g in gold["labels"].items():
            info = infos[int(lab)]
            assert info.nvox == g["nvox"], lab
            if g["nvox"] == 0:
                continue
            for ps, fb in ((1, 0), (1, 1)):
                e = g[f"p{ps}_b{fb}"]
                gv, gt, r = eng.meshify_label(d, info, 0.5, 0, ps, fb, 0)
                assert (len(gv), len(gt)) == (e["nverts"], e["ntris"]), (lab, ps, fb)
                assert topology_digest(gv, gt)[2] == e["digest"], (lab, ps, fb)
        # EVERY label against the unmodified reference's per-label mesh (tools/make_golden_big.py: whole-volume
        # binarisation + the reference's meshify(), -p1 -l0 -b0 iso 0.5): counts, topology + f32-exact positions, and
        # the four labels whose smoothed maximum stays below 0.5 so that the isolevel is reset (src/meshify.c:316-319)
        big = json.loads((GOLDEN / "golden_big.json").read_text())["atlas"]
        assert big["nlabel"] == gold["nlabel"] and big["nonempty"] == gold["nonempty"] == 365
        resets = []
        for info in infos[1:]:
            e = big["labels"][str(info.label)]
            assert info.nvox == e["nvox"], info.label
            if not info.nvox:
                continue
            gv, gt, r = eng.meshify_label(d, info, 0.5, 0, 1, 0, 0)
            assert (len(gv), len(gt)) == (e["nverts"], e["ntris"]), info.label
            assert topology_digest(gv, gt)[2] == e["digest"], info.label
            if r.iso_reset:
                resets.append(info.label)
        assert resets == big["resets"] == [127, 180, 188, 192]
    finally:
        d.free()


def test_atlas_all_labels_in_one_call(eng):
    """b2m_atlas_meshify_all(): the label loop inside the library, labels spread over worker contexts; every label equals
    the reference's recorded mesh (golden_big.json), the skipped ones are the reference's skipped ones"""
    from nii2mesh_b200 import synth
    big = json.loads((GOLDEN / "golden_big.json").read_text())["atlas"]
    vol, _ = synth.load_nifti(GOLDEN / "D99_atlas_v2.0_right.nii.gz")
    d = eng.upload(vol)
    try:
        for workers in (1, 8):
            res = eng.atlas_meshify_all(d, 0.5, 0, 1, 0, 0, workers=workers, fetch=True)
            assert len(res) == big["nlabel"]
            for lab, e in res.items():
                g = big["labels"][str(lab)]
                assert e["nvox"] == g["nvox"], lab
                if g["nvox"] == 0:
                    assert e["rc"] == -100, lab
                    continue
                assert e["rc"] == 0 and (e["nverts"], e["ntris"]) == (g["nverts"], g["ntris"]), lab
                assert topology_digest(e["verts"], e["tris"])[2] == g["digest"], lab
                assert bool(e["iso_reset"]) == (lab in big["resets"]), lab
    finally:
        d.free()
