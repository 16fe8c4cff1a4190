"""Canonical mesh comparison (TEST INFRASTRUCTURE).

Two meshes are "the same" when, after dropping vertices no triangle references (the reference keeps
orphans, /root/reference/src/meshify.c:113-114), there is a bijection between their vertices with
|a-b| <= 1e-5*max(1,|coord|) per coordinate, and the triangle lists — rewritten through the
bijection, each rotated so its smallest index comes first (winding preserved) and sorted
lexicographically — are identical (bit-exact topology).
"""
import hashlib

import numpy as np


def drop_orphans(verts, tris):
    verts = np.asarray(verts, np.float64)
    tris = np.asarray(tris, np.int64)
    used = np.zeros(len(verts), bool)
    used[tris.ravel()] = True
    remap = np.cumsum(used) - 1
    return verts[used], remap[tris]


def canon_faces(tris):
    t = np.asarray(tris, np.int64)
    k = np.argmin(t, axis=1)
    r = np.stack([np.take_along_axis(t, ((k + i) % 3)[:, None], 1)[:, 0] for i in range(3)], axis=1)
    order = np.lexsort((r[:, 2], r[:, 1], r[:, 0]))
    return r[order]


def match_vertices(va, vb, rtol=1e-5):
    """returns perm with va[i] ~ vb[perm[i]], or raises AssertionError."""
    assert len(va) == len(vb), f"vertex count differs: {len(va)} vs {len(vb)}"
    if len(va) == 0:
        return np.zeros(0, np.int64)
    # fast path: bit-identical coordinates
    a = np.ascontiguousarray(va).view([("", np.float64)] * 3).ravel()
    b = np.ascontiguousarray(vb).view([("", np.float64)] * 3).ravel()
    ia, ib = np.argsort(a, kind="stable"), np.argsort(b, kind="stable")
    if np.array_equal(a[ia], b[ib]) and len(np.unique(a)) == len(a):
        perm = np.empty(len(va), np.int64)
        perm[ia] = ib
        return perm
    from scipy.spatial import cKDTree
    tree = cKDTree(vb)
    d, j = tree.query(va, k=1)
    tol = rtol * np.maximum(1.0, np.abs(va).max(axis=1))
    assert np.all(np.abs(va - vb[j]).max(axis=1) <= tol), "vertex without a partner within tolerance"
    assert len(np.unique(j)) == len(j), "vertex matching is not a bijection"
    return j


def assert_same_mesh(verts_a, tris_a, verts_b, tris_b, rtol=1e-5):
    """a = implementation under test, b = oracle/reference."""
    va, ta = drop_orphans(verts_a, tris_a)
    vb, tb = drop_orphans(verts_b, tris_b)
    assert len(ta) == len(tb), f"triangle count differs: {len(ta)} vs {len(tb)}"
    perm = match_vertices(va, vb, rtol)
    fa = canon_faces(perm[ta])
    fb = canon_faces(tb)
    bad = np.nonzero((fa != fb).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} triangles differ after canonical sorting (first at {bad[:5]})"


def topology_digest(verts, tris, with_coords=True):
    """Order-independent digest of a mesh for golden files.  Vertices are ranked by their exact
    coordinates (the implementation reproduces the reference's positions bit-for-bit: same f32/f64
    operations in the same order), faces are canonicalised through that ranking.  with_coords=False
    hashes only the face list (pure topology).  Returns (n_used_verts, n_tris, sha256 hex)."""
    v, t = drop_orphans(verts, tris)
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0]))
    rank = np.empty(len(v), np.int64)
    rank[order] = np.arange(len(v))
    f = canon_faces(rank[t])
    h = hashlib.sha256()
    if with_coords:
        h.update(np.ascontiguousarray(v[order] + 0.0).tobytes())  # +0.0 folds -0.0 into +0.0
    h.update(np.ascontiguousarray(f).tobytes())
    return len(v), len(t), h.hexdigest()
