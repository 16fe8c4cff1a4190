"""Post-smooth: laplacian_smoothHC(), src/quadric.c:343-394.

CPU part: the oracle restatement (oracle/oracle.c: orc_laplacian_hc, a gather over an incidence list) against the
digests recorded from the unmodified reference (tests/golden/golden_post.json, tools/make_golden_post.py) and, where
oracle/_ref is present, against the live compiled reference on adversarial meshes.
GPU part (-m gpu): the CUDA path through the C ABI (b2m_laplacian_hc_host / _device, laplacian_smoothHC) against the
oracle and the same digests, bit for bit."""
import ctypes as C
import hashlib
import json

import numpy as np
import pytest

import cases
from conftest import GOLDEN, bits_differ
import oracle

GOLD = json.loads((GOLDEN / "golden_post.json").read_text())
RUNS = [(1, True), (3, True), (3, False), (10, True)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def orc():
    return oracle.Oracle()


@pytest.fixture(scope="module")
def meshes(orc):
    """the golden cases' meshes, rebuilt by the oracle's meshify (pinned to the reference's by digest)"""
    vols = cases.volumes()
    out = {}
    for key, g in GOLD.items():
        name, f = key.split("/")
        omc, p, l, b = int(f[1]), int(f[3]), int(f[5]), int(f[7])
        vol, iso = vols[name]
        m = orc.meshify(vol, iso, omc, p, l, b)
        assert m["rc"] == 0 and sha(m["verts"].tobytes() + m["tris"].tobytes()) == g["mesh"]
        out[key] = (m["verts"], m["tris"])
    return out


def adversarial():
    """open patch + unreferenced vertices, triangles with repeated indices, random index soup, no triangles, fan"""
    n = 12
    gv = np.array([[x, y, np.sin(x * y)] for y in range(n) for x in range(n)], float)
    gt = []
    for y in range(n - 1):
        for x in range(n - 1):
            a = y * n + x
            gt += [[a, a + 1, a + n], [a + 1, a + n + 1, a + n]]
    gt = np.array(gt, np.int32)
    gv2 = np.vstack([gv, [[100, 100, 100], [5, 5, 5]]])
    rng = np.random.default_rng(3)
    rv = rng.normal(size=(200, 3))
    fan_v = np.vstack([[[0, 0, 1.0]], [[np.cos(a), np.sin(a), 0] for a in np.linspace(0, 2 * np.pi, 400, endpoint=False)]])
    fan_t = np.array([[0, 1 + k, 1 + (k + 1) % 400] for k in range(400)], np.int32)  # one vertex with 400 triangles
    return {
        "patch": (gv2, gt),
        "repeated": (gv2, np.vstack([gt, [[0, 0, 5], [7, 7, 7], [3, 9, 3]]]).astype(np.int32)),
        "soup": (rv, rng.integers(0, 200, size=(500, 3)).astype(np.int32)),
        "notris": (rv, np.zeros((0, 3), np.int32)),
        "fan": (fan_v, fan_t),
        "one": (rv[:3], np.array([[0, 1, 2]], np.int32)),
    }


# ---- CPU ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", list(GOLD))
def test_oracle_matches_reference_digests(orc, meshes, key):
    v, t = meshes[key]
    for it, lock in RUNS:
        assert sha(orc.laplacian_hc(v, t, it, lock_edges=lock)) == GOLD[key][f"iter{it}_lock{int(lock)}"], (key, it, lock)


@pytest.mark.skipif(not oracle.ref_available("lewiner"), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", list(adversarial()))
def test_oracle_matches_live_reference(orc, name):
    ref = oracle.Ref("lewiner")
    v, t = adversarial()[name]
    for it, lock, a, b in [(0, True, 0.1, 0.5), (1, True, 0.1, 0.5), (2, False, 0.1, 0.5), (4, True, 0.3, 0.2)]:
        x = orc.laplacian_hc(v, t, it, alpha=a, beta=b, lock_edges=lock)
        y = ref.laplacian_hc(v, t, it, alpha=a, beta=b, lock_edges=lock)
        assert bits_differ(x, y) == 0, (name, it, lock)


def test_border_vertices_stay(orc):
    v, t = adversarial()["patch"]
    s = orc.laplacian_hc(v, t, 3, lock_edges=True)
    n = 12
    rim = [y * n + x for y in range(n) for x in range(n) if x in (0, n - 1) or y in (0, n - 1)]
    assert np.array_equal(s[rim], v[rim]) and np.array_equal(s[-2:], v[-2:])  # rim and unreferenced vertices untouched
    inner = [y * n + x for y in range(2, n - 2) for x in range(2, n - 2)]
    assert not np.array_equal(s[inner], v[inner])


# ---- GPU ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("key", list(GOLD))
def test_gpu_matches_oracle_and_digests(eng, orc, meshes, key):
    v, t = meshes[key]
    for it, lock in RUNS:
        g = eng.laplacian_hc(v, t, it, lock_edges=lock)
        assert sha(g) == GOLD[key][f"iter{it}_lock{int(lock)}"], (key, it, lock)
        assert bits_differ(g, orc.laplacian_hc(v, t, it, lock_edges=lock)) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(adversarial()))
def test_gpu_adversarial_meshes(eng, orc, name):
    v, t = adversarial()[name]
    for it, lock, a, b in [(0, True, 0.1, 0.5), (1, True, 0.1, 0.5), (2, False, 0.1, 0.5), (4, True, 0.3, 0.2)]:
        assert bits_differ(eng.laplacian_hc(v, t, it, alpha=a, beta=b, lock_edges=lock),
                           orc.laplacian_hc(v, t, it, alpha=a, beta=b, lock_edges=lock)) == 0, (name, it, lock)


@pytest.mark.gpu
def test_gpu_device_resident_mesh_and_reference_prototype(eng, orc):
    """meshify on the device, post-smooth the device mesh in place, fetch: equals meshify -> host laplacian_smoothHC()"""
    vol, iso = cases.volumes()["gyroid96"]
    d = eng.upload(vol)
    _, _, r = eng.meshify_device(d, iso, 0, 1, 1, 0, fetch=False)
    v0, t0 = eng.fetch(r)
    eng.laplacian_hc_result(r, 5)
    v1, t1 = eng.fetch(r)
    d.free()
    assert np.array_equal(t0, t1)
    assert bits_differ(v1, orc.laplacian_hc(v0, t0, 5)) == 0
    # the reference's prototype (include/quadric.h), verts updated in place
    v2 = v0.copy()
    eng.lib.laplacian_smoothHC(v2.ctypes.data, t0.ctypes.data, len(v2), len(t0), 0.1, 0.5, 5, True)
    assert bits_differ(v2, v1) == 0


@pytest.mark.gpu
def test_gpu_rejects_bad_indices(eng):
    from nii2mesh_b200 import lib
    v = np.zeros((4, 3))
    with pytest.raises(lib.B2MError, match="outside"):
        eng.laplacian_hc(v, np.array([[0, 1, 4]], np.int32), 1)
    with pytest.raises(lib.B2MError, match="outside"):
        eng.laplacian_hc(v, np.array([[0, -1, 2]], np.int32), 1)
    assert np.array_equal(eng.laplacian_hc(v, np.array([[0, 1, 2]], np.int32), 0), v)
