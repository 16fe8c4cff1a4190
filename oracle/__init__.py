"""oracle — TEST INFRASTRUCTURE ONLY.

CPU checker for the nii2mesh voxel->mesh hot path.  Two things live here:

* ``Oracle``  — ctypes front-end to ``oracle/oracle.c``, the plain-C restatement of the reference
  algorithm (each C function cites the reference file:line it follows).
* ``Ref``     — ctypes front-end to the UNMODIFIED reference compiled by ``oracle/build_ref.sh`` into
  ``oracle/_ref/libref_{lewiner,classic}.so`` (git-ignored; present here and shipped to the GPU box,
  absent from a fresh clone).  Used to pin the restatement and as the CPU baseline.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product (``nii2mesh_b200``) never does.
"""
import ctypes as C
import os
import sys
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
LIB = HERE / "liboracle.so"


class vec3d(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class vec3i(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("z", C.c_int)]


def build(force=False):
    """Compile oracle.c -> liboracle.so, and (if the reference sources are present) oracle/_ref."""
    src = HERE / "oracle.c"
    inc = HERE.parent / "nii2mesh_b200" / "csrc" / "mc_tables.inc"
    stale = (not LIB.exists()) or LIB.stat().st_mtime < max(src.stat().st_mtime, inc.stat().st_mtime)
    if force or stale:
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", str(src), "-o", str(LIB), "-lm"])
    if Path("/root/reference/src/meshify.c").exists() and (
            force or not (REF_DIR / "libref_lewiner.so").exists()):
        subprocess.check_call(["bash", str(HERE / "build_ref.sh")], stdout=subprocess.DEVNULL)
    return LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _take_mesh(libc_free, pp, pt, nv, nt):
    v = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_double)), shape=(nv, 3)).copy() if nv else np.zeros((0, 3))
    t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(nt, 3)).copy() if nt else np.zeros((0, 3), np.int32)
    libc_free(pp)
    libc_free(pt)
    return v, t


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]


def _to_malloc(arr):
    """copy a numpy array into a malloc() block (the reference free()s / realloc()s what it is given)."""
    arr = np.ascontiguousarray(arr)
    p = _libc.malloc(max(arr.nbytes, 8))
    C.memmove(p, arr.ctypes.data, arr.nbytes)
    return p


class Oracle:
    """The C restatement.  Volumes are numpy float32 arrays indexed [z, y, x] (x fastest)."""

    def __init__(self):
        build()
        self.lib = C.CDLL(str(LIB))
        L = self.lib
        L.orc_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_bwlabel.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.orc_cc_label.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 4
        L.orc_dilate25.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_front.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                C.c_void_p]
        L.orc_mc_lewiner.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.c_int, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_mc_classic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_weld.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]
        L.orc_degenerate.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_meshify.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]

    def set_threshold(self, vol, dark_medium_bright_123):
        """restatement of setThreshold() (src/isolevel.c:245-277): -i d / m / b"""
        v = _f32(vol)
        self.lib.orc_set_threshold.restype = C.c_float
        self.lib.orc_set_threshold.argtypes = [C.c_void_p, C.c_int, C.c_int]
        return float(self.lib.orc_set_threshold(v.ctypes.data, v.size, int(dark_medium_bright_123)))

    def smooth(self, vol):
        v = _f32(vol).copy()
        nz, ny, nx = v.shape
        self.lib.orc_smooth(v.ctypes.data, nx, ny, nz)
        return v

    def cc_label(self, bw, conn):
        bw = np.ascontiguousarray(bw, dtype=np.uint8)
        nz, ny, nx = bw.shape
        lab = np.zeros(bw.shape, np.uint32)
        n = self.lib.orc_cc_label(bw.ctypes.data, lab.ctypes.data, nx, ny, nz, conn)
        return lab, n

    def bwlabel(self, mask, conn=18, only_largest=True, fill_bubbles=False):
        m = _f32(mask).copy()
        nz, ny, nx = m.shape
        self.lib.orc_bwlabel(m.ctypes.data, conn, nx, ny, nz, int(only_largest), int(fill_bubbles))
        return m

    def dilate25(self, mask):
        m = _f32(mask).copy()
        nz, ny, nx = m.shape
        self.lib.orc_dilate25(m.ctypes.data, nx, ny, nz)
        return m

    def front(self, vol, iso, pre_smooth, only_largest, fill_bubbles):
        """returns dict(img=composed volume, iso, lo, hi, mn, mx, mask, rc)"""
        v = _f32(vol).copy()
        nz, ny, nx = v.shape
        isoc = C.c_float(iso)
        lo = (C.c_int * 3)()
        hi = (C.c_int * 3)()
        mn = C.c_float()
        mx = C.c_float()
        mask = np.zeros(v.shape, np.float32)
        rc = self.lib.orc_front(v.ctypes.data, nx, ny, nz, C.byref(isoc), int(pre_smooth), int(only_largest),
                                int(fill_bubbles), lo, hi, C.byref(mn), C.byref(mx), mask.ctypes.data)
        return dict(img=v, iso=isoc.value, lo=list(lo), hi=list(hi), mn=mn.value, mx=mx.value, mask=mask, rc=rc)

    def mc(self, img, lo, hi, iso, original_mc=0, backend=0):
        v = _f32(img)
        nz, ny, nx = v.shape
        lo_ = (C.c_int * 3)(*lo)
        hi_ = (C.c_int * 3)(*hi)
        pp, pt, nv, nt = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        if backend == 1:
            rc = self.lib.orc_mc_classic(v.ctypes.data, nx, ny, nz, lo_, hi_, iso, C.byref(pp), C.byref(pt),
                                         C.byref(nv), C.byref(nt))
        else:
            rc = self.lib.orc_mc_lewiner(v.ctypes.data, nx, ny, nz, lo_, hi_, int(original_mc), iso, C.byref(pp),
                                         C.byref(pt), C.byref(nv), C.byref(nt))
        if rc:
            return None
        return _take_mesh(_libc.free, pp, pt, nv.value, nt.value)

    def weld(self, verts, tris):
        nv, nt = len(verts), len(tris)
        pp = C.c_void_p(_to_malloc(np.asarray(verts, np.float64)))
        t = np.ascontiguousarray(tris, dtype=np.int32).copy()
        n2 = self.lib.orc_weld(C.byref(pp), t.ctypes.data, nv, nt)
        v = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_double)), shape=(n2, 3)).copy()
        _libc.free(pp)
        return v, t

    def degenerate(self, verts, tris):
        v = np.ascontiguousarray(verts, dtype=np.float64)
        t = np.ascontiguousarray(tris, dtype=np.int32).copy()
        n2 = self.lib.orc_degenerate(v.ctypes.data, t.ctypes.data, len(t))
        return t[:n2].copy()

    def laplacian_hc(self, verts, tris, iters, alpha=0.1, beta=0.5, lock_edges=True):
        v = np.ascontiguousarray(verts, dtype=np.float64).copy()
        t = np.ascontiguousarray(tris, dtype=np.int32)
        self.lib.orc_laplacian_hc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        self.lib.orc_laplacian_hc.restype = None
        self.lib.orc_laplacian_hc(v.ctypes.data, t.ctypes.data, len(v), len(t), alpha, beta, int(iters), int(lock_edges))
        return v

    def meshify(self, vol, iso, original_mc=0, pre_smooth=True, only_largest=True, fill_bubbles=False, backend=0):
        """returns dict(verts, tris, pre_nv, pre_nt, iso, rc); vol is not modified."""
        v = _f32(vol).copy()
        nz, ny, nx = v.shape
        pp, pt, nv, nt = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        pnv, pnt, isou = C.c_int(), C.c_int(), C.c_float()
        rc = self.lib.orc_meshify(v.ctypes.data, nx, ny, nz, int(original_mc), iso, C.byref(pt), C.byref(pp),
                                  C.byref(nt), C.byref(nv), int(pre_smooth), int(only_largest), int(fill_bubbles),
                                  int(backend), C.byref(pnv), C.byref(pnt), C.byref(isou))
        if rc:
            return dict(rc=rc)
        verts, tris = _take_mesh(_libc.free, pp, pt, nv.value, nt.value)
        return dict(rc=0, verts=verts, tris=tris, pre_nv=pnv.value, pre_nt=pnt.value, iso=isou.value)


def ref_available(flavour="lewiner"):
    return (REF_DIR / f"libref_{flavour}.so").exists()


class Ref:
    """The unmodified reference, compiled from /root/reference/src by oracle/build_ref.sh.
    flavour 'lewiner' (make lewiner) or 'classic' (default make, -DUSE_CLASSIC_CUBES)."""

    def __init__(self, flavour="lewiner"):
        path = REF_DIR / f"libref_{flavour}.so"
        if not path.exists():
            raise FileNotFoundError(path)
        self.flavour = flavour
        self.lib = C.CDLL(str(path), mode=os.RTLD_LOCAL)
        L = self.lib
        L.meshify.argtypes = [C.c_void_p, C.POINTER(C.c_short), C.c_int, C.c_float, C.POINTER(C.c_void_p),
                              C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_bool, C.c_bool,
                              C.c_bool, C.c_bool]
        L.quick_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.bwlabel.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.c_bool, C.c_bool]
        L.dilate.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_bool]
        L.marchingCubes.argtypes = [C.c_void_p, C.POINTER(C.c_short), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                    C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]
        L.unify_vertices.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_bool]
        L.remove_degenerate_triangles.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_bool]

    def set_threshold(self, vol, dark_medium_bright_123):
        """setThreshold() (src/isolevel.c:245-277): -i d / m / b"""
        v = _f32(vol)
        self.lib.setThreshold.restype = C.c_float
        self.lib.setThreshold.argtypes = [C.c_void_p, C.c_int, C.c_int]
        return float(self.lib.setThreshold(v.ctypes.data, v.size, int(dark_medium_bright_123)))

    def smooth(self, vol):
        v = _f32(vol).copy()
        nz, ny, nx = v.shape
        self.lib.quick_smooth(v.ctypes.data, nx, ny, nz)
        return v

    def bwlabel(self, mask, conn=18, only_largest=True, fill_bubbles=False):
        m = _f32(mask).copy()
        nz, ny, nx = m.shape
        dim = (C.c_size_t * 3)(nx, ny, nz)
        self.lib.bwlabel(m.ctypes.data, conn, dim, bool(only_largest), bool(fill_bubbles))
        return m

    def dilate(self, mask):
        m = _f32(mask).copy()
        nz, ny, nx = m.shape
        dim = (C.c_size_t * 3)(nx, ny, nz)
        self.lib.dilate(m.ctypes.data, dim, True)
        return m

    def mc(self, img, lo, hi, iso, original_mc=0):
        v = _f32(img)
        nz, ny, nx = v.shape
        dim = (C.c_short * 3)(nx, ny, nz)
        pp, pt, nv, nt = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        rc = self.lib.marchingCubes(v.ctypes.data, dim, (C.c_int * 3)(*lo), (C.c_int * 3)(*hi), int(original_mc), iso,
                                    C.byref(pp), C.byref(pt), C.byref(nv), C.byref(nt))
        if rc:
            return None
        return _take_mesh(_libc.free, pp, pt, nv.value, nt.value)

    def weld(self, verts, tris):
        nv, nt = len(verts), len(tris)
        pp = C.c_void_p(_to_malloc(np.asarray(verts, np.float64)))
        t = np.ascontiguousarray(tris, dtype=np.int32).copy()
        n2 = self.lib.unify_vertices(C.byref(pp), t.ctypes.data, nv, nt, False)
        v = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_double)), shape=(n2, 3)).copy()
        _libc.free(pp)
        return v, t

    def degenerate(self, verts, tris):
        v = np.ascontiguousarray(verts, dtype=np.float64)
        pt = C.c_void_p(_to_malloc(np.asarray(tris, np.int32)))
        n2 = self.lib.remove_degenerate_triangles(v.ctypes.data, C.byref(pt), len(tris), False)
        t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(n2, 3)).copy()
        _libc.free(pt)
        return t

    def laplacian_hc(self, verts, tris, iters, alpha=0.1, beta=0.5, lock_edges=True):
        """laplacian_smoothHC(), src/quadric.c:343-394 (what nii2mesh -s <iters> runs after meshify + apply_sform)"""
        v = np.ascontiguousarray(verts, dtype=np.float64).copy()
        t = np.ascontiguousarray(tris, dtype=np.int32)
        self.lib.laplacian_smoothHC.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_bool]
        self.lib.laplacian_smoothHC.restype = None
        self.lib.laplacian_smoothHC(v.ctypes.data, t.ctypes.data, len(v), len(t), alpha, beta, int(iters), bool(lock_edges))
        return v

    def meshify(self, vol, iso, original_mc=0, pre_smooth=True, only_largest=True, fill_bubbles=False,
                return_img=False, counts=False):
        """counts=True: run verbose with the process's stdout (fd 1) redirected to a temporary file and parse the
        reference's own diagnostics (src/meshify.c:84,104,166,318) -> pre_nv, pre_nt, iso_reset.  NOT thread-safe
        (fd 1 is process-wide): callers that want parallelism use processes."""
        v = _f32(vol).copy()
        nz, ny, nx = v.shape
        dim = (C.c_short * 3)(nx, ny, nz)
        pp, pt, nv, nt = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        log = ""
        if counts:
            import tempfile
            _libc.fflush(None)
            sys.stdout.flush()
            tmp = tempfile.TemporaryFile(mode="w+b")
            saved = os.dup(1)
            os.dup2(tmp.fileno(), 1)
        try:
            rc = self.lib.meshify(v.ctypes.data, dim, int(original_mc), iso, C.byref(pt), C.byref(pp), C.byref(nt),
                                  C.byref(nv), bool(pre_smooth), bool(only_largest), bool(fill_bubbles), bool(counts))
        finally:
            if counts:
                _libc.fflush(None)
                os.dup2(saved, 1)
                os.close(saved)
                tmp.seek(0)
                log = tmp.read().decode(errors="replace")
                tmp.close()
        if rc:
            return dict(rc=rc, log=log)
        verts, tris = _take_mesh(_libc.free, pp, pt, nv.value, nt.value)
        out = dict(rc=0, verts=verts, tris=tris)
        if return_img:
            out["img"] = v
        if counts:
            import re
            m = re.search(r"vertex welding (\d+) -> (\d+)", log)
            out["pre_nv"] = int(m.group(1)) if m else len(verts)
            m = re.search(r"remove degenerate triangles (\d+) -> (\d+)", log)
            out["pre_nt"] = int(m.group(1)) if m else len(tris)
            out["iso_reset"] = "Suggested isolevel out of range" in log
            out["log"] = log
        return out
