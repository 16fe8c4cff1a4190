"""GPU tests of the z-slab path through the C ABI (b2m_meshify_slab).

A volume cut into 2..4 z-slabs — halo exchange for the smooth and marching cubes, connected
components merged across the slab faces, globally numbered vertices, seam vertices welded — must
give the SAME arrays as the single-volume path (which the parity suite pins to the reference):
vertices bit for bit, triangles index for index.  Here every slab is a b2m_ctx driven by its own
host thread on cuda:0 (the single-process transport, b2m_comm_create_local); the NCCL transport runs
the same code above a different exchange layer and is covered by tests/test_slabs_nccl.py + bench.py.
"""
import numpy as np
import pytest

import cases
from oracle.canon import assert_same_mesh

pytestmark = pytest.mark.gpu

VOLS = cases.volumes()


@pytest.fixture(scope="module")
def groups():
    from nii2mesh_b200 import lib
    gs = {w: lib.LocalSlabGroup(w) for w in (2, 3, 4)}
    yield gs
    for g in gs.values():
        g.close()


def _cuts(nz, world, uneven):
    if not uneven:
        return [round(i * nz / world) for i in range(world + 1)]
    # uneven slabs, each >= 4 planes
    rng = np.random.default_rng(nz * 7 + world)
    while True:
        inner = sorted(rng.choice(np.arange(4, nz - 3), world - 1, replace=False).tolist())
        c = [0] + inner + [nz]
        if min(b - a for a, b in zip(c, c[1:])) >= 4:
            return c


SLAB_VOLS = ["sphere40", "blobs", "blobs_w33", "blobs2", "gyroid96", "sphere64", "bet", "isoreset"]


@pytest.mark.parametrize("name", SLAB_VOLS)
@pytest.mark.parametrize("world", [2, 3, 4])
def test_slabs_equal_single_volume(eng, groups, name, world):
    vol, iso = VOLS[name]
    nz = vol.shape[0]
    if nz < 4 * world:
        pytest.skip("volume too thin for this many slabs")
    d = eng.upload(vol)
    try:
        for k, (backend, omc, ps, ol, fb) in enumerate(cases.flag_sets(name)):
            sv, st, sr = eng.meshify_device(d, iso, omc, ps, ol, fb, backend)
            cuts = _cuts(nz, world, uneven=bool(k & 1))
            gv, gt, rs = groups[world].meshify(vol, cuts, iso, original_mc=omc, pre_smooth=ps, only_largest=ol,
                                               fill_bubbles=fb, backend=backend)
            tag = f"{name} w{world} cuts{cuts} b{backend} o{omc} p{ps} l{ol} f{fb}"
            r = rs[0].r
            assert (r.nverts, r.ntris, r.pre_nverts, r.pre_ntris) == (sr.nverts, sr.ntris, sr.pre_nverts, sr.pre_ntris), tag
            assert (list(r.lo), list(r.hi), r.iso_used, r.vmin, r.vmax) == (list(sr.lo), list(sr.hi), sr.iso_used, sr.vmin, sr.vmax), tag
            assert np.array_equal(gt, st), tag
            assert np.array_equal(gv.view(np.uint64), sv.view(np.uint64)), tag
    finally:
        d.free()


def test_slabs_components_snake_through_seams(eng, groups):
    """a component that crosses every seam several times, islands and bubbles that straddle seams, equal-size
    clusters (tie -> earliest first voxel) on different slabs"""
    nz, ny, nx = 48, 40, 70
    v = np.full((nz, ny, nx), -1.0, np.float32)
    # a serpentine: vertical bars joined alternately at the top and at the bottom
    for i, x in enumerate(range(4, 60, 8)):
        v[4:44, 10:14, x:x + 3] = 1.0
        z = 41 if i % 2 == 0 else 4
        v[z:z + 3, 10:14, x:x + 11] = 1.0
    # two clusters of exactly equal size on different slabs + a bigger one made of the serpentine
    v[6:10, 25:29, 5:9] = 1.0
    v[36:40, 25:29, 5:9] = 1.0
    # a hollow box straddling the middle seam (a bubble to fill) and one open to the volume face
    v[18:30, 22:34, 40:52] = 1.0
    v[21:27, 25:31, 43:49] = -1.0
    v[0:8, 30:38, 58:66] = 1.0
    v[0:5, 32:36, 60:64] = -1.0
    v += np.random.default_rng(3).normal(0, 0.01, v.shape).astype(np.float32)
    d = eng.upload(v)
    try:
        for world in (2, 3, 4):
            for ps, ol, fb in ((0, 1, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1)):
                sv, st, sr = eng.meshify_device(d, 0.0, 0, ps, ol, fb, 0)
                cuts = _cuts(nz, world, False)
                gv, gt, rs = groups[world].meshify(v, cuts, 0.0, original_mc=0, pre_smooth=ps, only_largest=ol,
                                                   fill_bubbles=fb, backend=0)
                tag = f"w{world} p{ps} l{ol} b{fb}"
                assert np.array_equal(gt, st), tag
                assert np.array_equal(gv.view(np.uint64), sv.view(np.uint64)), tag
    finally:
        d.free()


def test_slabs_vs_oracle(eng, orc, groups):
    """and directly against the CPU oracle (not only against our own single-volume path)"""
    vol, iso = VOLS["blobs2"]
    for backend, omc, ps, ol, fb in ((0, 0, 1, 1, 1), (1, 0, 1, 1, 0), (0, 1, 0, 0, 0)):
        o = orc.meshify(vol, iso, omc, ps, ol, fb, backend)
        gv, gt, rs = groups[3].meshify(vol, _cuts(vol.shape[0], 3, True), iso, original_mc=omc, pre_smooth=ps,
                                       only_largest=ol, fill_bubbles=fb, backend=backend)
        assert (len(gv), len(gt)) == (len(o["verts"]), len(o["tris"]))
        assert_same_mesh(gv, gt, o["verts"], o["tris"], 1e-5)


def test_slabs_bad_geometry_fails_loudly(eng, groups):
    from nii2mesh_b200 import lib
    vol, iso = VOLS["sphere40"]
    with pytest.raises(lib.B2MError):
        groups[2].meshify(vol, [0, 38, 40], iso)  # a 2-plane slab
    # the group is usable again afterwards
    gv, gt, _ = groups[2].meshify(vol, [0, 20, 40], iso)
    assert len(gv) > 0 and len(gt) > 0


def _snake_volume():
    nz, ny, nx = 48, 40, 70
    v = np.full((nz, ny, nx), -1.0, np.float32)
    for i, x in enumerate(range(4, 60, 8)):
        v[4:44, 10:14, x:x + 3] = 1.0
        z = 41 if i % 2 == 0 else 4
        v[z:z + 3, 10:14, x:x + 11] = 1.0
    v[6:10, 25:29, 5:9] = 1.0
    v[36:40, 25:29, 5:9] = 1.0
    v[18:30, 22:34, 40:52] = 1.0
    v[21:27, 25:31, 43:49] = -1.0
    v += np.random.default_rng(3).normal(0, 0.01, v.shape).astype(np.float32)
    return v


@pytest.mark.parametrize("caps", [("1", "1"), ("3", "100000"), ("100000", "2")])
def test_slabs_seam_block_overflow_takes_the_slow_path(eng, groups, monkeypatch, caps):
    """the fast seam merge ships one fixed-size block of seam roots / pairs per rank; when a rank has more (noise
    volumes), every rank sees the overflow flag in the gathered headers, untags its roots and the sorted-list path
    takes over.  Forced here with tiny capacities; the result must not change."""
    monkeypatch.setenv("B2M_SEAM_ENT_CAP", caps[0])
    monkeypatch.setenv("B2M_SEAM_PAIR_CAP", caps[1])
    for vol, iso in ((_snake_volume(), 0.0), VOLS["blobs"], VOLS["sphere40"]):
        d = eng.upload(vol)
        try:
            for world in (2, 4):
                if vol.shape[0] < 4 * world:
                    continue
                for ps, ol, fb in ((0, 1, 1), (1, 1, 0), (0, 0, 1)):
                    sv, st, sr = eng.meshify_device(d, iso, 0, ps, ol, fb, 0)
                    gv, gt, rs = groups[world].meshify(vol, _cuts(vol.shape[0], world, False), iso, original_mc=0, pre_smooth=ps,
                                                       only_largest=ol, fill_bubbles=fb, backend=0)
                    assert np.array_equal(gt, st) and np.array_equal(gv.view(np.uint64), sv.view(np.uint64)), (caps, world, ps, ol, fb)
        finally:
            d.free()


def test_slabs_unaligned_slab_pointers(eng, groups):
    """slabs handed in as d_volume + z0 * nx * ny of ONE device volume whose plane size is odd: only 4-byte aligned
    (the range reduction of the -p 0 path read float4 from such pointers)"""
    from nii2mesh_b200 import lib
    vol, iso = VOLS["blobs"]          # 30 x 37 x 41: 1517 voxels per plane
    nz, ny, nx = vol.shape
    d = eng.upload(vol)
    try:
        for world, cuts in ((2, [0, 13, nz]), (3, [0, 9, 19, nz])):   # odd plane counts: the slabs start 4-byte aligned only
            views = [lib.DeviceVolume(eng, (cuts[i + 1] - cuts[i], ny, nx), d.ptr.value + cuts[i] * ny * nx * 4) for i in range(world)]
            assert any(v.ptr % 16 for v in views)
            for ps, ol, fb in ((0, 0, 0), (0, 1, 1), (1, 1, 0)):
                sv, st, _ = eng.meshify_device(d, iso, 0, ps, ol, fb, 0)
                gv, gt, _ = groups[world].meshify_device(views, vol.shape, cuts, iso, original_mc=0, pre_smooth=ps,
                                                         only_largest=ol, fill_bubbles=fb, backend=0)
                assert np.array_equal(gt, st) and np.array_equal(gv.view(np.uint64), sv.view(np.uint64)), (world, ps, ol, fb)
    finally:
        d.free()


@pytest.mark.timeout(600)
def test_slab_host_local_group_large_mesh(eng):
    """b2m_meshify_slab_host() driven by the host threads of ONE process (local transport) with output blocks beyond the
    pre-fault threshold (64 MiB per rank): ADVICE r1's deadlock - a rank kept the process-wide copy pool while it waited
    in a collective for a rank that was waiting for the pool.  G512 in two slabs; the assembled mesh is the reference's
    (counts + topology digest recorded from the unmodified reference, tests/golden/golden_big.json)."""
    import json
    import threading
    from conftest import GOLDEN
    from nii2mesh_b200 import lib, synth, slabs
    from oracle.canon import topology_digest
    n = 512
    vol = np.tile(synth.gyroid_tile(128), (n // 128,) * 3)
    grp = lib.LocalSlabGroup(2)
    cuts = [0, 256, 512]
    out, err = [None, None], [None, None]

    def work(i):
        try:
            r, v, t = grp.engs[i].meshify_slab_host(grp.comms[i], vol[cuts[i]:cuts[i + 1]], vol.shape, cuts[i], 0.0, 0, 1, 1, 1, 0)
            out[i] = slabs.part_of(r, v, t)
        except Exception as ex:  # noqa: BLE001
            err[i] = ex
    try:
        for rep in range(2):   # the second call pre-faults its blocks from the first call's totals, before the H2D
            th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not any(err), err
            assert min((p["nv_edge"] + p["nv_cent"]) * 24 + p["ntris_local"] * 12 for p in out) > (64 << 20)
        V, T = slabs.assemble(out)
    finally:
        grp.close()
    g = json.loads((GOLDEN / "golden_big.json").read_text())["gyroid"]["512"]
    assert (len(V), len(T)) == (g["nverts"], g["ntris"]) == (5802752, 11606149)
    assert topology_digest(V, T)[2] == g["digest"]
