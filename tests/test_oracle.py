"""CPU tests of the oracle (oracle/oracle.c): pinned against (a) golden digests produced by the
unmodified compiled reference (tests/golden/golden.json, tools/make_golden.py) and (b) the compiled
reference itself, live, whenever oracle/_ref is present.  No GPU needed."""
import hashlib
import json

import numpy as np
import pytest

import cases
import surfaces
from conftest import GOLDEN, bits_differ
from oracle.canon import assert_same_mesh, topology_digest

GOLD = json.loads((GOLDEN / "golden.json").read_text())
VOLS = cases.volumes()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- golden vectors (reference outputs recorded in the build container) ---------------------------
@pytest.mark.parametrize("name", list(VOLS))
def test_smooth_golden(orc, name):
    assert sha(orc.smooth(VOLS[name][0])) == GOLD["smooth"][name]


@pytest.mark.parametrize("name", list(VOLS))
def test_bwlabel_dilate_golden(orc, name):
    vol, iso = VOLS[name]
    mask = (vol >= np.float32(iso)).astype(np.float32)
    for ol, fb in ((1, 0), (0, 1), (1, 1)):
        assert sha(orc.bwlabel(mask, 18, ol, fb) != 0) == GOLD["bwlabel"][f"{name}/l{ol}b{fb}"], (ol, fb)
    assert sha(orc.dilate25(mask) != 0) == GOLD["dilate"][name]


@pytest.mark.parametrize("k", range(10))
def test_selftest_surfaces_known_answers(orc, k):
    """the reference's own MC self test: 10 analytic 60^3 surfaces (src/MarchingCubes.c:1282-1362)"""
    v, t = orc.mc(surfaces.surface(k), [0, 0, 0], [59, 59, 59], 0.0, 0, 0)
    assert (len(v), len(t)) == surfaces.KNOWN[k]
    assert topology_digest(v, t)[2] == GOLD["surfaces"][str(k)]["digest"]


def test_selftest_surface_original_mc(orc):
    v, t = orc.mc(surfaces.surface(7), [0, 0, 0], [59, 59, 59], 0.0, 1, 0)
    assert (len(v), len(t)) == surfaces.KNOWN_ORIGINAL[7]
    assert topology_digest(v, t)[2] == GOLD["surfaces"]["7_original"]["digest"]


@pytest.mark.parametrize("name", list(VOLS))
def test_meshify_golden(orc, name):
    vol, iso = VOLS[name]
    for backend, omc, ps, ol, fb in cases.flag_sets(name):
        key = f"{name}/backend{backend}_o{omc}_p{ps}_l{ol}_b{fb}"
        g = GOLD["meshify"][key]
        o = orc.meshify(vol, iso, omc, ps, ol, fb, backend)
        assert o["rc"] == g["rc"], key
        if g["rc"]:
            continue
        assert (len(o["verts"]), len(o["tris"])) == (g["nverts"], g["ntris"]), key
        assert topology_digest(o["verts"], o["tris"])[2] == g["digest"], key
        f = orc.front(vol, iso, ps, ol, fb)
        assert sha(f["img"]) == GOLD["front"][key], key


def test_no_variability_fails(orc):
    assert orc.meshify(cases.flat_volume(), 1.0)["rc"] != 0


# ---- live against the compiled reference --------------------------------------------------------
SMALL = [n for n in VOLS if VOLS[n][0].size < 400000]


@pytest.mark.parametrize("name", SMALL)
def test_stages_vs_reference(orc, ref_lewiner, ref_classic, name):
    vol, iso = VOLS[name]
    assert bits_differ(orc.smooth(vol), ref_lewiner.smooth(vol)) == 0
    mask = (vol >= np.float32(iso)).astype(np.float32)
    for ol, fb in ((1, 0), (0, 1), (1, 1), (0, 0)):
        assert np.array_equal(orc.bwlabel(mask, 18, ol, fb), ref_lewiner.bwlabel(mask, 18, ol, fb)), (ol, fb)
    assert np.array_equal(orc.dilate25(mask), ref_lewiner.dilate(mask))
    f = orc.front(vol, iso, 0, 0, 0)
    if f["rc"]:
        return
    for omc in (0, 1):
        ov, ot = orc.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, 0)
        rv, rt = ref_lewiner.mc(f["img"], f["lo"], f["hi"], f["iso"], omc)
        assert np.array_equal(ot, rt) and np.array_equal(ov, rv), f"lewiner o{omc}"
    ov, ot = orc.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, 1)
    rv, rt = ref_classic.mc(f["img"], f["lo"], f["hi"], f["iso"], 0)
    assert np.array_equal(ot, rt) and np.array_equal(ov, rv), "classic"
    # weld + degenerate removal on the classic soup (most merges)
    wv, wt = orc.weld(ov, ot)
    xv, xt = ref_classic.weld(rv, rt)
    assert np.array_equal(wv, xv) and np.array_equal(wt, xt)
    assert np.array_equal(orc.degenerate(wv, wt), ref_classic.degenerate(xv, xt))


@pytest.mark.parametrize("name", ["sphere40", "blobs", "thin4", "isoreset"])
def test_meshify_vs_reference(orc, ref_lewiner, ref_classic, name):
    vol, iso = VOLS[name]
    for backend, omc, ps, ol, fb in cases.flag_sets(name):
        R = ref_classic if backend else ref_lewiner
        r = R.meshify(vol, iso, omc, ps, ol, fb)
        o = orc.meshify(vol, iso, omc, ps, ol, fb, backend)
        assert o["rc"] == r["rc"]
        if r["rc"] == 0:
            assert np.array_equal(o["verts"], r["verts"]) and np.array_equal(o["tris"], r["tris"])


def _narrow_volumes():
    rng = np.random.default_rng(21)
    return {"nx1": rng.standard_normal((12, 10, 1)).astype(np.float32), "ny1": rng.standard_normal((12, 1, 10)).astype(np.float32),
            "nx1_ny1": rng.standard_normal((20, 1, 1)).astype(np.float32), "nz1": rng.standard_normal((1, 9, 11)).astype(np.float32)}


def test_narrow_volumes_vs_reference(orc, ref_lewiner):
    """bwlabelCore() refuses dim[0] < 2 or dim[1] < 2 and leaves the mask as thresholded (src/bwlabel.c:434-437):
    -l / -b become no-ops there, while a single-plane volume (nz == 1) is labelled normally"""
    for name, vol in _narrow_volumes().items():
        mask = (vol >= np.float32(0.1)).astype(np.float32)
        for ol, fb in ((1, 0), (0, 1), (1, 1)):
            assert np.array_equal(orc.bwlabel(mask, 18, ol, fb), ref_lewiner.bwlabel(mask, 18, ol, fb)), (name, ol, fb)
        for ps, ol, fb in ((0, 1, 0), (0, 1, 1), (1, 1, 1)):
            r = ref_lewiner.meshify(vol, 0.1, 0, ps, ol, fb)
            o = orc.meshify(vol, 0.1, 0, ps, ol, fb, 0)
            assert o["rc"] == r["rc"], (name, ps, ol, fb)
            if r["rc"] == 0:
                assert np.array_equal(o["verts"], r["verts"]) and np.array_equal(o["tris"], r["tris"])


def test_weld_adversarial_vs_reference(orc, ref_lewiner):
    """near-duplicate vertices around the 1e-5 tolerance, incl. chains (later heads steal)"""
    rng = np.random.default_rng(11)
    for trial in range(20):
        base = rng.uniform(0, 40, (60, 3))
        pts = [base]
        for s in (2e-6, 6e-6, 9.9e-6, 1.2e-5):
            pts.append(base[rng.integers(0, 60, 25)] + rng.normal(0, s, (25, 3)))
        v = np.concatenate(pts)
        v = v[rng.permutation(len(v))]
        t = rng.integers(0, len(v), (300, 3)).astype(np.int32)
        wv, wt = orc.weld(v, t)
        xv, xt = ref_lewiner.weld(v, t)
        assert np.array_equal(wv, xv) and np.array_equal(wt, xt), trial
        assert np.array_equal(orc.degenerate(wv, wt), ref_lewiner.degenerate(xv, xt)), trial


def test_canon_detects_differences(orc):
    vol, iso = VOLS["sphere24"]
    o = orc.meshify(vol, iso, 0, 1, 1, 0, 0)
    v, t = o["verts"], o["tris"]
    perm = np.random.default_rng(0).permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    assert_same_mesh(v[perm], inv[t][::-1], v, t)  # relabelled + reordered: same mesh
    bad = t.copy()
    bad[0] = bad[0][[1, 0, 2]]  # flipped winding
    with pytest.raises(AssertionError):
        assert_same_mesh(v, bad, v, t)


def _isolevel_cases():
    rng = np.random.default_rng(1)
    out = {k: v[0] for k, v in cases.volumes().items()}
    out["binary"] = (rng.random((20, 20, 20)) > 0.7).astype(np.float32)
    out["const"] = np.full((5, 5, 5), 2.0, np.float32)
    out["tiny"] = rng.standard_normal((3, 4, 5)).astype(np.float32)                      # < 100 voxels: full range
    out["nan"] = np.where(rng.random((16, 16, 16)) > 0.9, np.nan, rng.standard_normal((16, 16, 16))).astype(np.float32)
    out["spike"] = np.concatenate([np.zeros(5000), np.ones(5000), [1000.0]]).astype(np.float32).reshape(1, 1, -1)
    return out


def test_isolevel_restatement_equals_reference(orc, ref_lewiner):
    """-i d / m / b: setThreshold() (src/isolevel.c:245-277): robust range + Otsu; BASELINE config 1 uses bet 'medium'"""
    for name, vol in _isolevel_cases().items():
        for mode in (1, 2, 3):
            a, b = orc.set_threshold(vol, mode), ref_lewiner.set_threshold(vol, mode)
            assert a == b or (np.isnan(a) and np.isnan(b)), (name, mode, a, b)
    assert abs(orc.set_threshold(cases.volumes()["bet"][0], 2) - 67.729) < 1e-3


# ---- the GPU clean-up's pre-filter as a property of the reference's own output -------------
def _removed_mask(before, after):
    """order-preserving compaction (src/meshify.c:147-166): which rows of `before` are missing from `after`"""
    rem = np.ones(len(before), bool)
    j = 0
    for i in range(len(before)):
        if j < len(after) and np.array_equal(before[i], after[j]):
            rem[i] = False
            j += 1
    assert j == len(after)
    return rem


@pytest.mark.parametrize("name,backend,flags", [
    ("sphere40", cases.LEWINER, (0, 0, 0, 0)), ("sphere64", cases.LEWINER, (0, 1, 1, 0)), ("blobs", cases.LEWINER, (0, 0, 0, 0)),
    ("blobs2", cases.LEWINER, (1, 0, 0, 0)), ("bet", cases.LEWINER, (0, 1, 1, 0)), ("gyroid160", cases.LEWINER, (0, 1, 1, 1)),
    ("sphere40", cases.CLASSIC, (0, 0, 0, 0)), ("blobs", cases.CLASSIC, (0, 1, 1, 1))])
def test_degenerate_triangles_touch_a_grid_corner(orc, name, backend, flags):
    """libb2m runs the reference's needle test (src/meshify.c:113-145) only for triangles with a vertex within 1/128 of
    a grid corner along its cube edge, or a vertex that is not an edge vertex (Lewiner centroid vertices): every triangle
    the test removes must be such a triangle - and by a wide margin (the closest removed triangle is reported)."""
    vol, iso = VOLS[name]
    omc, p, l, b = flags
    f = orc.front(vol, iso, p, l, b)
    v, t = orc.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, backend)
    v2, t2 = orc.weld(v, t)
    kept = orc.degenerate(v2, t2)
    rem = _removed_mask(t2, kept)
    frac = np.abs(v2 - np.rint(v2))                       # distance of every coordinate to the grid
    on_grid = frac == 0.0
    edge_vertex = on_grid.sum(axis=1) >= 2                # two integer coordinates: a vertex on a cube edge
    free = np.where(on_grid, 0.0, frac).max(axis=1)       # its free coordinate's distance to the nearest corner
    slow = ~edge_vertex | (free < 1.0 / 128)
    tri_slow = slow[t2].any(axis=1)
    assert rem.sum() == len(t2) - len(kept)
    assert not (rem & ~tri_slow).any(), f"{(rem & ~tri_slow).sum()} removed triangles without a near-corner vertex"
    if rem.any():
        closest = np.where(edge_vertex[t2], free[t2], 0.0).min(axis=1)[rem].max()
        assert closest < 1e-3, closest                    # 150x inside the 1/128 band (area >= 0.3 d^2 argument)
