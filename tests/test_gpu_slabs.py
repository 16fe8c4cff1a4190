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
