"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
libb2m.so (include/b2m.h, include/meshify.h); results are compared with the CPU oracle on the same
inputs and with the golden digests recorded from the unmodified reference.

Bar: integer/index work (masks, bbox, triangle topology) bit-exact; vertex positions within 1e-5
relative (north_star) — in fact the kernels reproduce the reference's f32/f64 operations in the
same order, so positions are compared bit-for-bit wherever the emission order is the reference's.
"""
import ctypes as C
import hashlib
import json

import numpy as np
import pytest

import cases
import surfaces
from conftest import GOLDEN, bits_differ
from oracle.canon import assert_same_mesh, topology_digest

pytestmark = pytest.mark.gpu

GOLD = json.loads((GOLDEN / "golden.json").read_text())
VOLS = cases.volumes()
POS_RTOL = 1e-5  # north_star: vertex positions within 1e-5 relative


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", list(VOLS))
def test_smooth_bit_exact(eng, orc, name):
    vol = VOLS[name][0]
    a = eng.smooth(vol)
    assert bits_differ(a, orc.smooth(vol)) == 0
    assert sha(a) == GOLD["smooth"][name]


def test_smooth_unsafe_inputs_bit_exact(eng, orc):
    """denormals, -0, huge values: the kernel's FP64-adder rounding shortcut must hand these planes to
    the real f64->f32 conversions (and still match the reference bit for bit)"""
    rng = np.random.default_rng(0)
    v = rng.standard_normal((20, 33, 40)).astype(np.float32)
    v[3:6, 4:9, 5:30] = 1e-42
    v[7, 7, 7] = -0.0
    v[10:12, 10:20, 3:9] = 3e38
    v[15, 5, 5] = 1e-39
    v[2, 2, 2:12] = 0.0
    assert bits_differ(eng.smooth(v), orc.smooth(v)) == 0
    z = np.zeros((9, 12, 64), np.float32)
    z[4, 6, 30] = 1.0
    assert bits_differ(eng.smooth(z), orc.smooth(z)) == 0
    assert bits_differ(eng.smooth(-z), orc.smooth(-z)) == 0


@pytest.mark.parametrize("name", list(VOLS))
def test_front_masks_bbox_bit_exact(eng, orc, name):
    """smooth -> range -> isolevel sanity -> CC (largest / bubbles) -> dilate -> darken -> bbox"""
    vol, iso = VOLS[name]
    for ps, ol, fb in ((0, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1), (0, 1, 1), (0, 1, 0)):
        a = eng.front(vol, iso, ps, ol, fb)
        b = orc.front(vol, iso, ps, ol, fb)
        tag = f"{name} p{ps} l{ol} b{fb}"
        assert (a["lo"], a["hi"]) == (b["lo"], b["hi"]), tag
        assert (a["iso"], a["mn"], a["mx"]) == (b["iso"], b["mn"], b["mx"]), tag
        if ol or fb:
            assert np.array_equal(a["mask"] != 0, b["mask"] != 0), tag
        assert bits_differ(a["img"], b["img"]) == 0, tag


@pytest.mark.parametrize("name", list(VOLS))
def test_front_golden(eng, name):
    vol, iso = VOLS[name]
    for backend, omc, ps, ol, fb in cases.flag_sets(name):
        if backend:
            continue
        key = f"{name}/backend{backend}_o{omc}_p{ps}_l{ol}_b{fb}"
        if key in GOLD["front"]:
            assert sha(eng.front(vol, iso, ps, ol, fb)["img"]) == GOLD["front"][key], key


@pytest.mark.parametrize("name", list(VOLS))
def test_marching_cubes_arrays(eng, orc, name):
    """Lewiner: vertex and triangle ARRAYS identical to the reference's emission order;
    classic: edge-keyed mesh == the reference's soup after its own weld."""
    vol, iso = VOLS[name]
    for ps, ol in ((1, 1), (0, 0)):
        f = orc.front(vol, iso, ps, ol, 0)
        for omc in (0, 1):
            ov, ot = orc.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, 0)
            gv, gt, _ = eng.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, 0)
            assert np.array_equal(gt, ot), f"{name} p{ps} o{omc}"
            assert np.array_equal(gv, ov), f"{name} p{ps} o{omc}"
        sv, st = orc.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, 1)
        wv, wt = orc.weld(sv, st)
        gv, gt, _ = eng.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, 1)
        cv, ct = eng.weld(gv, gt)  # resolves the rare corner merges exactly like the reference
        assert_same_mesh(cv, ct, wv, orc.degenerate(wv, wt), POS_RTOL)


@pytest.mark.parametrize("k", range(10))
def test_selftest_surfaces(eng, k):
    v, t, _ = eng.mc(surfaces.surface(k), [0, 0, 0], [59, 59, 59], 0.0, 0, 0)
    assert (len(v), len(t)) == surfaces.KNOWN[k]
    assert topology_digest(v, t)[2] == GOLD["surfaces"][str(k)]["digest"]
    if k == 7:
        v, t, _ = eng.mc(surfaces.surface(7), [0, 0, 0], [59, 59, 59], 0.0, 1, 0)
        assert (len(v), len(t)) == surfaces.KNOWN_ORIGINAL[7]
        assert topology_digest(v, t)[2] == GOLD["surfaces"]["7_original"]["digest"]


@pytest.mark.parametrize("name", [n for n in VOLS if VOLS[n][0].size < 300000])
def test_weld_hook(eng, orc, name):
    vol, iso = VOLS[name]
    f = orc.front(vol, iso, 0, 0, 0)
    for backend in (0, 1):
        ov, ot = orc.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, backend)
        wv, wt = orc.weld(ov, ot)
        wt = orc.degenerate(wv, wt)
        gv, gt = eng.weld(ov, ot)
        assert len(gv) == len(wv)
        assert_same_mesh(gv, gt, wv, wt, POS_RTOL)


def test_weld_adversarial(eng, orc):
    rng = np.random.default_rng(11)
    for trial in range(10):
        base = rng.uniform(0, 40, (60, 3))
        pts = [base]
        for s in (2e-6, 6e-6, 9.9e-6, 1.2e-5):
            pts.append(base[rng.integers(0, 60, 25)] + rng.normal(0, s, (25, 3)))
        v = np.concatenate(pts)
        v = v[rng.permutation(len(v))]
        t = rng.integers(0, len(v), (300, 3)).astype(np.int32)
        wv, wt = orc.weld(v, t)
        wt = orc.degenerate(wv, wt)
        gv, gt = eng.weld(v, t)
        assert len(gv) == len(wv) and len(gt) == len(wt), trial
        assert_same_mesh(gv, gt, wv, wt, POS_RTOL)


@pytest.mark.parametrize("name", list(VOLS))
def test_meshify_vs_oracle_and_golden(eng, orc, name):
    """the whole path, host buffers in / host mesh out (b2m_meshify_host = what meshify() calls)"""
    vol, iso = VOLS[name]
    for backend, omc, ps, ol, fb in cases.flag_sets(name):
        key = f"{name}/backend{backend}_o{omc}_p{ps}_l{ol}_b{fb}"
        g = GOLD["meshify"][key]
        gv, gt, r = eng.meshify(vol, iso, omc, ps, ol, fb, backend)
        assert (len(gv), len(gt)) == (g["nverts"], g["ntris"]), key
        nu, nt, faces = topology_digest(gv, gt, with_coords=False)
        assert nu == g["nused"], key
        if backend == 0:
            # Lewiner: bit-exact positions + topology against the reference's recorded digest
            assert topology_digest(gv, gt)[2] == g["digest"], key
        if vol.size < 600000 or backend == 1:
            # classic: FP64 positions may differ in the last bits (which soup copy of an edge vertex the
            # reference's weld keeps), so compare through the tolerance matcher (1e-5 relative)
            o = orc.meshify(vol, iso, omc, ps, ol, fb, backend)
            assert (r.pre_nverts, r.pre_ntris) == (o["pre_nv"], o["pre_nt"]), key
            assert_same_mesh(gv, gt, o["verts"], o["tris"], POS_RTOL)


def _c_meshify(libb2m):
    libb2m.meshify.argtypes = [C.c_void_p, C.POINTER(C.c_short), C.c_int, C.c_float, C.POINTER(C.c_void_p),
                               C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_bool, C.c_bool,
                               C.c_bool, C.c_bool]
    return libb2m.meshify


def test_reference_signature_meshify(libb2m, orc, bet):
    """meshify() exactly as nii2() calls it (src/nii2mesh.c:326): short dims, malloc'd outputs"""
    vol, _ = bet
    fn = _c_meshify(libb2m)
    img = vol.copy()
    nz, ny, nx = img.shape
    dim = (C.c_short * 3)(nx, ny, nz)
    pt, pp, nt, nv = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    rc = fn(img.ctypes.data, dim, 0, 67.729, C.byref(pt), C.byref(pp), C.byref(nt), C.byref(nv), True, True, False, False)
    assert rc == 0
    g = GOLD["meshify"]["bet/backend0_o0_p1_l1_b0"]
    assert (nv.value, nt.value) == (g["nverts"], g["ntris"]) == (172304, 344400)
    v = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_double)), shape=(nv.value, 3)).copy()
    t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(nt.value, 3)).copy()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(pp)  # plain malloc() blocks owned by the caller (src/nii2mesh.c:353-354)
    libc.free(pt)
    assert topology_digest(v, t)[2] == g["digest"]
    # classic back-end through the runtime switch that replaces -DUSE_CLASSIC_CUBES
    libb2m.b2m_set_default_backend(1)
    try:
        rc = fn(img.ctypes.data, dim, 0, 67.729, C.byref(pt), C.byref(pp), C.byref(nt), C.byref(nv), True, True, False,
                False)
        assert rc == 0
        g = GOLD["meshify"]["bet/backend1_o0_p1_l1_b0"]
        assert (nv.value, nt.value) == (g["nverts"], g["ntris"])
        libc.free(pp)
        libc.free(pt)
    finally:
        libb2m.b2m_set_default_backend(0)


def test_failure_semantics(eng, libb2m):
    from nii2mesh_b200 import lib
    with pytest.raises(lib.MeshifyFailure):  # "No variability in image intensity" -> EXIT_FAILURE
        eng.meshify(cases.flat_volume(), 1.0)
    fn = _c_meshify(libb2m)
    img = cases.flat_volume()
    dim = (C.c_short * 3)(10, 9, 8)
    pt, pp, nt, nv = C.c_void_p(), C.c_void_p(), C.c_int(-7), C.c_int(-7)
    rc = fn(img.ctypes.data, dim, 0, 1.0, C.byref(pt), C.byref(pp), C.byref(nt), C.byref(nv), True, True, False, False)
    assert rc == 1 and not pt.value and not pp.value and nt.value == -7  # outputs untouched on failure
    # isolevel outside the intensity range is reset to mid-range (src/meshify.c:316-319)
    v, t, r = eng.meshify(VOLS["isoreset"][0], 1.0e6, 0, 1, 1, 0)
    assert r.iso_reset == 1 and r.iso_used == np.float32(0.5 * (np.float64(np.float32(r.vmin + r.vmax))))


def test_input_not_modified_and_deterministic(eng):
    vol, iso = VOLS["blobs2"]
    keep = vol.copy()
    a = eng.meshify(vol, iso, 0, 1, 1, 1)
    b = eng.meshify(vol, iso, 0, 1, 1, 1)
    assert np.array_equal(vol, keep)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# ---- full-size, size-independent properties -----------------------------------------------------
def _gyroid_counts(n):  # pre-weld counts of the G family are cubic in tiles/axis
    return 79416 * n ** 3 + 45456 * n ** 2 - 1410 * n, 158848 * n ** 3 + 90912 * n ** 2 - 2832 * n


@pytest.mark.parametrize("n,post", [(256, (814168, 1628383)), (512, (5802752, 11606149)),
                                    (1024, (43545074, 87096797))])
def test_gyroid_full_size_known_counts(eng, n, post):
    """BASELINE config 3 family (Lewiner -p1 -l1 -b1): the reference's recorded counts
     before and after weld/cleanup, at 256^3, 512^3 and the full 1024^3."""
    from nii2mesh_b200 import synth
    d = eng.tiled_volume(synth.gyroid_tile(128), (n, n, n))
    try:
        _, _, r = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
        assert (r.pre_nverts, r.pre_ntris) == _gyroid_counts(n // 128)
        assert (r.nverts, r.ntris) == post
        v, t, r2 = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=(n <= 512))
        assert (r2.nverts, r2.ntris) == post  # idempotent: the device volume is not modified
        if v is not None:
            # closed-surface bookkeeping that does not depend on size: every index in range and used
            assert t.min() >= 0 and t.max() < len(v)
            assert len(np.unique(t)) >= len(v) - r.ndegenerate
            # topology + bit-exact positions against the unmodified reference's mesh of the same volume (recorded by
            # tools/make_golden_big.py: 4 s / 30 s of reference CPU time)
            g = json.loads((GOLDEN / "golden_big.json").read_text())["gyroid"][str(n)]
            assert (len(v), len(t)) == (g["nverts"], g["ntris"])
            assert topology_digest(v, t)[2] == g["digest"]
    finally:
        d.free()


def test_sphere512_classic_known_counts(eng):
    """BASELINE config 2: S512 noisy sphere, original marching cubes, -p 0 -l 0."""
    from nii2mesh_b200 import synth
    vol = synth.noisy_sphere(512)
    d = eng.upload(vol)
    try:
        _, _, r = eng.meshify_device(d, 0.0, 1, 0, 0, 0, 1, fetch=False)   # classic build
        assert (r.pre_nverts, r.pre_ntris, r.nverts, r.ntris) == (16032504, 5344168, 2791829, 5343010)
        _, _, r = eng.meshify_device(d, 0.0, 1, 0, 0, 0, 0, fetch=False)   # Lewiner build, -o 1
        assert (r.pre_nverts, r.pre_ntris, r.nverts, r.ntris) == (2791860, 5331520, 2791807, 5330338)
    finally:
        d.free()


def test_cc_tile_root_list_overflow_fallback(eng, orc, monkeypatch):
    """the compact tile-root list of the CC passes overflows on noise volumes: the full-scan variants must give the
    same masks (forced here with a zero capacity)"""
    monkeypatch.setenv("B2M_CC_LIST_CAP", "0")
    for name in ("blobs", "sphere40", "blobs2"):
        vol, iso = VOLS[name]
        for ps, ol, fb in ((0, 1, 1), (1, 1, 0), (0, 0, 1)):
            a = eng.front(vol, iso, ps, ol, fb)
            b = orc.front(vol, iso, ps, ol, fb)
            assert np.array_equal(a["mask"] != 0, b["mask"] != 0), (name, ps, ol, fb)
            assert bits_differ(a["img"], b["img"]) == 0, (name, ps, ol, fb)


def test_isolevel_selection(eng, orc, libb2m):
    """-i d / m / b through b2m_isolevel_device / _host and the reference-named setThreshold(): the reference's float"""
    import test_oracle
    for name, vol in test_oracle._isolevel_cases().items():
        d = eng.upload(vol)
        try:
            for mode in (1, 2, 3):
                want = orc.set_threshold(vol, mode)
                got_d, got_h = eng.isolevel(d, mode), eng.isolevel(vol, mode)
                assert got_d == want or (np.isnan(got_d) and np.isnan(want)), (name, mode, got_d, want)
                assert got_h == got_d or (np.isnan(got_h) and np.isnan(got_d)), (name, mode)
        finally:
            d.free()
    bet = np.ascontiguousarray(VOLS["bet"][0])
    iso = libb2m.setThreshold(bet.ctypes.data, bet.size, 2)
    assert iso == orc.set_threshold(bet, 2) and abs(iso - 67.729) < 1e-3   # BASELINE config 1: "default medium isolevel"


def test_ingest_and_sform_on_device(eng, orc, libb2m):
    """load_nii's voxel conversion (src/nii2mesh.c:155-172) and apply_sform (src/meshify.c:1021-1045) on the GPU:
    raw u8 / i16 / u16 / f32 voxels in, world-space mesh out, equal to convert-on-host + meshify + host apply_sform"""
    rng = np.random.default_rng(9)
    base = VOLS["blobs2"][0]
    srow = [[-0.7, 0.01, 0.0, 90.5], [0.02, 0.6, -0.03, -126.25], [0.0, 0.05, 0.9, -72.0]]   # negative determinant proxy: winding flips
    for dt, slope, inter in ((np.uint8, 0.025, -3.0), (np.int16, 0.001, 0.5), (np.uint16, 0.0, 0.0), (np.float32, 1.5, -0.25)):
        if dt == np.float32:
            raw = base.astype(np.float32)
        else:
            info = np.iinfo(dt)
            span = base.max() - base.min()
            raw = ((base - base.min()) / span * min(info.max, 4000)).astype(dt)
        s = np.float32(slope if slope != 0.0 else 1.0)
        vol = (raw.astype(np.float32) * s) + np.float32(inter)          # f32 product, f32 sum
        d = eng.ingest(raw, slope, inter)
        try:
            assert bits_differ(d.to_host(), vol) == 0, dt
        finally:
            d.free()
        iso = float(np.float32(vol.mean()))
        for backend in (0, 1):
            hv, ht, _ = eng.meshify(vol, iso, 0, 1, 1, 1, backend)
            # the reference's apply_sform on the host mesh (our meshify_host.c copy is pinned to it in test_abi)
            v2 = np.ascontiguousarray(hv.copy())
            t2 = np.ascontiguousarray(ht.copy())
            libb2m.apply_sform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
            libb2m.apply_sform.restype = None
            rows = [(C.c_float * 4)(*r) for r in srow]
            libb2m.apply_sform(t2.ctypes.data, v2.ctypes.data, len(t2), len(v2), rows[0], rows[1], rows[2])
            gv, gt, r = eng.meshify_raw(raw, iso, slope, inter, srow, 0, 1, 1, 1, backend)
            assert np.array_equal(gt, t2) and np.array_equal(gv.view(np.uint64), v2.view(np.uint64)), (dt, backend)
            assert not np.array_equal(gt, ht)   # the winding really flipped
    # without srow: plain voxel coordinates, identical to meshify()
    raw = (np.clip(base, -2, 2) * 1000).astype(np.int16)
    vol = raw.astype(np.float32) * np.float32(0.001)
    hv, ht, _ = eng.meshify(vol, 0.2, 0, 1, 1, 0, 0)
    gv, gt, _ = eng.meshify_raw(raw, 0.2, 0.001, 0.0, None, 0, 1, 1, 0, 0)
    assert np.array_equal(gt, ht) and np.array_equal(gv.view(np.uint64), hv.view(np.uint64))


def test_repeated_host_calls_large_pageable_volume(eng):
    """two meshify() calls in a row on a pageable host volume whose mesh exceeds the pre-fault threshold (64 MiB): the
    second call starts pre-faulting its output blocks BEFORE the pageable H2D copy, which needs the same copy pool
    (regression test for a self-deadlock), and both give the same mesh"""
    from nii2mesh_b200 import synth
    vol = synth.gyroid(384)
    a = eng.meshify(vol, 0.0, 0, 1, 1, 1, 0)
    b = eng.meshify(vol, 0.0, 0, 1, 1, 1, 0)
    assert a[2].nverts * 24 + a[2].ntris * 12 > (64 << 20)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))


def test_narrow_volumes(eng, orc):
    """bwlabelCore() refuses volumes narrower than 2 voxels in x or y (src/bwlabel.c:434-437): -l / -b are no-ops there"""
    import test_oracle
    for name, vol in test_oracle._narrow_volumes().items():
        for ps, ol, fb in ((0, 1, 0), (0, 1, 1), (0, 0, 1), (1, 1, 1)):
            a, b = eng.front(vol, 0.1, ps, ol, fb), orc.front(vol, 0.1, ps, ol, fb)
            assert np.array_equal(a["mask"] != 0, b["mask"] != 0), (name, ps, ol, fb)
            assert bits_differ(a["img"], b["img"]) == 0 and a["lo"] == b["lo"] and a["hi"] == b["hi"], (name, ps, ol, fb)


@pytest.mark.parametrize("shape", [(40, 50, 96), (19, 24, 128), (23, 70, 160), (12, 9, 32)])
def test_smooth_writes_the_threshold_bit_rows(eng, orc, shape, monkeypatch):
    """B2M_SMOOTH_BITS=1 (opt-in: correct but slower on B200), nx % 32 == 0: k_smooth3 writes
    fg / bg / mb for the requested isolevel and k_threshold is skipped (csrc/smooth.cu, csrc/pipeline.cu b2m_front_run).  Partial tiles in x and y, an isolevel one ulp above a smoothed voxel (that voxel
    is outside for `>= iso` but inside for marching cubes' `v - iso > -FLT_EPSILON`), an isolevel the range check
    replaces (the speculated rows are then discarded), both table sets."""
    monkeypatch.setenv("B2M_SMOOTH_BITS", "1")
    rng = np.random.default_rng(shape[2] * 7 + shape[0])
    vol = rng.standard_normal(shape).astype(np.float32)
    vol[shape[0] // 3: shape[0] // 2, 2:-2, 3:-3] += 2.0  # one big bright block + speckle
    S = orc.smooth(vol)
    cand = S[(np.abs(S) > 0.25) & (np.abs(S) < 0.5)]
    near = np.nextafter(np.float32(cand[len(cand) // 2]), np.float32(np.inf))
    for iso in (0.3, float(near), 1.0, 1e9, -1e9):
        for ps, ol, fb in ((1, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1)):
            a, b = eng.front(vol, iso, ps, ol, fb), orc.front(vol, iso, ps, ol, fb)
            tag = (shape, iso, ps, ol, fb)
            assert (a["iso"], a["lo"], a["hi"]) == (b["iso"], b["lo"], b["hi"]), tag
            if ol or fb:
                assert np.array_equal(a["mask"] != 0, b["mask"] != 0), tag
            assert bits_differ(a["img"], b["img"]) == 0, tag
        for backend in (0, 1):
            gv, gt, r = eng.meshify(vol, iso, 0, 1, 1, 1, backend)
            o = orc.meshify(vol, iso, 0, 1, 1, 1, backend)
            assert (r.pre_nverts, r.pre_ntris) == (o["pre_nv"], o["pre_nt"]), (shape, iso, backend)
            assert_same_mesh(gv, gt, o["verts"], o["tris"], POS_RTOL)
        # the path under test really ran: no k_threshold launch unless the range check replaced the isolevel
        eng.set_profile(True)
        try:
            d = eng.upload(vol)
            eng.meshify_device(d, iso, 0, 1, 1, 1, 0, fetch=False)
            names = [k for k, _ in eng.kernel_times()]
        finally:
            eng.set_profile(False)
        assert ("threshold" in names) == (abs(iso) > 1e8), (shape, iso, names)


def test_pinned_input_overlapped_h2d(eng):
    """a pinned host volume >= 256 MiB goes up in z-chunks on a second stream while the smooth follows the transfer
    (b2m_meshify_host); the mesh equals the device-resident path bit for bit"""
    from nii2mesh_b200 import lib, synth
    n = 448   # 343 MiB
    L = eng.lib
    hp = C.c_void_p()
    eng._chk(L.b2m_host_alloc(C.byref(hp), n ** 3 * 4))
    try:
        hvol = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(n, n, n))
        hvol[...] = synth.gyroid(n)
        o = lib.Opts(0.0, 0, 1, 1, 1, 0, 0)
        r = lib.Result()
        pv, pt = C.c_void_p(), C.c_void_p()
        eng._chk(L.b2m_meshify_host(eng.ctx, hp, (C.c_int64 * 3)(n, n, n), C.byref(o), C.byref(pv), C.byref(pt), C.byref(r)))
        v = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(r.nverts, 3)).copy()
        t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(r.ntris, 3)).copy()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(pv)
        libc.free(pt)
        d = eng.upload(hvol)
        try:
            v2, t2, _ = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0)
        finally:
            d.free()
        assert np.array_equal(t, t2) and np.array_equal(v.view(np.uint64), v2.view(np.uint64))
    finally:
        L.b2m_host_free(hp)


def test_registered_destination_d2h_path():
    """few cores per rank (several ranks on one node): the output blocks are registered piecewise and the DMA engine
    writes straight into them instead of going through the pinned ring.  Forced here with B2M_D2H_REGISTER=1 (read once
    per process, hence the child process); G512: two blocks of 139 MB; the mesh must equal the recorded reference digest"""
    import subprocess
    import sys
    code = (
        "import json, sys, numpy as np\n"
        "sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "from nii2mesh_b200 import lib, synth\n"
        "from oracle.canon import topology_digest\n"
        "eng = lib.Engine(0)\n"
        "vol = np.tile(synth.gyroid_tile(128), (4, 4, 4))\n"
        "for rep in range(2):\n"
        "    v, t, r = eng.meshify(vol, 0.0, 0, 1, 1, 1, 0)\n"
        "g = json.load(open('tests/golden/golden_big.json'))['gyroid']['512']\n"
        "assert (len(v), len(t)) == (g['nverts'], g['ntris']), (len(v), len(t))\n"
        "assert topology_digest(v, t)[2] == g['digest']\n"
        "print('REGISTERED_D2H_OK', r.d2h_ms)\n")
    from conftest import ROOT
    import os
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, B2M_D2H_REGISTER="1"))
    assert p.returncode == 0 and "REGISTERED_D2H_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
