"""ctypes binding of libb2m.so — the host-side mirror of the reference interface in Python.

`Engine.meshify()` has the argument meaning of the reference's meshify()
(/root/reference/src/meshify.h:9): volume, isolevel, originalMC, preSmooth, onlyLargest,
fillBubbles; it returns (verts[np,3] f64, tris[nt,3] i32) like the reference's malloc'd vec3d/vec3i
arrays.  There is no CPU fallback: constructing an Engine without a CUDA device raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIBPATH = Path(os.environ.get("B2M_LIBPATH") or HERE / "libb2m.so")  # B2M_LIBPATH: an experiment build (build.py -D...)

BACKEND_LEWINER, BACKEND_CLASSIC = 0, 1
STAGES = ("smooth", "range", "cc", "compose", "mc", "weld", "degen", "total")


class Opts(C.Structure):
    _fields_ = [("isolevel", C.c_float), ("original_mc", C.c_int), ("pre_smooth", C.c_int), ("only_largest", C.c_int),
                ("fill_bubbles", C.c_int), ("backend", C.c_int), ("verbose", C.c_int)]


class Result(C.Structure):
    _fields_ = [("d_verts", C.c_void_p), ("d_tris", C.c_void_p), ("nverts", C.c_int), ("ntris", C.c_int),
                ("pre_nverts", C.c_int), ("pre_ntris", C.c_int), ("nmerged", C.c_int), ("ndegenerate", C.c_int),
                ("iso_used", C.c_float), ("vmin", C.c_float), ("vmax", C.c_float), ("lo", C.c_int * 3),
                ("hi", C.c_int * 3), ("iso_reset", C.c_int), ("ms", C.c_float * 8), ("launches", C.c_uint64),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("d2h_bytes", C.c_uint64)]

    def times(self):
        return {k: float(self.ms[i]) for i, k in enumerate(STAGES)}


class SlabResult(C.Structure):
    """b2m_slab_result: one rank's part of the mesh of a z-slab run (include/b2m.h)"""
    _fields_ = [("r", Result), ("d_verts", C.c_void_p), ("d_tris", C.c_void_p), ("nv_edge", C.c_int),
                ("nv_cent", C.c_int), ("nv_extra", C.c_int), ("ntris_local", C.c_int), ("v_edge_off", C.c_int64),
                ("v_cent_off", C.c_int64), ("v_extra_off", C.c_int64), ("tri_off", C.c_int64)]


class LabelInfo(C.Structure):
    """b2m_label_info: voxel count and bounding box of one atlas label"""
    _fields_ = [("label", C.c_int), ("nvox", C.c_longlong), ("lo", C.c_int * 3), ("hi", C.c_int * 3)]


class LabelMesh(C.Structure):
    """b2m_label_mesh: one entry of b2m_atlas_meshify_all()"""
    _fields_ = [("label", C.c_int), ("rc", C.c_int), ("nvox", C.c_longlong), ("nverts", C.c_int), ("ntris", C.c_int),
                ("verts", C.c_void_p), ("tris", C.c_void_p), ("r", Result)]


class B2MError(RuntimeError):
    pass


class MeshifyFailure(B2MError):
    """the reference's EXIT_FAILURE (no variability, empty mesh, < 3 vertices)"""


_lib = None


def load():
    """dlopen libb2m.so (must have been built: python -m nii2mesh_b200.build). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIBPATH.exists():
        raise B2MError(f"{LIBPATH} is missing: build it with `python -m nii2mesh_b200.build` (no CPU fallback exists)")
    L = C.CDLL(str(LIBPATH))
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.b2m_create.argtypes = [C.POINTER(vp), C.c_int]
    L.b2m_destroy.argtypes = [vp]
    L.b2m_destroy.restype = None
    L.b2m_last_error.restype = C.c_char_p
    L.b2m_version.restype = C.c_char_p
    L.b2m_set_default_backend.argtypes = [C.c_int]
    L.b2m_set_default_backend.restype = None
    L.b2m_dev_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.b2m_ctx_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.b2m_dev_free.argtypes = [vp]
    L.b2m_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.b2m_host_free.argtypes = [vp]
    L.b2m_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.b2m_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.b2m_sync.argtypes = [vp]
    L.b2m_flush_l2.argtypes = [vp]
    L.b2m_set_profile.argtypes = [vp, C.c_int]
    L.b2m_profile_count.argtypes = [vp]
    L.b2m_profile_entry.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.b2m_timer_start.argtypes = [vp]
    L.b2m_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.b2m_tile_volume.argtypes = [vp, vp, i64p, vp, i64p, C.c_int64]
    L.b2m_meshify_device.argtypes = [vp, vp, i64p, C.POINTER(Opts), C.POINTER(Result)]
    L.b2m_meshify_host.argtypes = [vp, vp, i64p, C.POINTER(Opts), C.POINTER(vp), C.POINTER(vp), C.POINTER(Result)]
    L.b2m_fetch_mesh.argtypes = [vp, C.POINTER(Result), vp, vp]
    L.b2m_atlas_scan.argtypes = [vp, vp, i64p, C.POINTER(C.c_int), C.POINTER(C.POINTER(LabelInfo))]
    L.b2m_atlas_free.argtypes = [C.POINTER(LabelInfo)]
    L.b2m_atlas_free.restype = None
    L.b2m_meshify_label_device.argtypes = [vp, vp, i64p, C.POINTER(LabelInfo), C.POINTER(Opts), C.POINTER(Result)]
    L.b2m_atlas_meshify_all.argtypes = [vp, vp, i64p, C.POINTER(Opts), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.POINTER(LabelMesh))]
    L.b2m_atlas_meshes_free.argtypes = [C.POINTER(LabelMesh), C.c_int]
    L.b2m_atlas_meshes_free.restype = None
    L.b2m_isolevel_device.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(C.c_float)]
    L.b2m_isolevel_host.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(C.c_float)]
    L.setThreshold.argtypes = [vp, C.c_int, C.c_int]
    L.setThreshold.restype = C.c_float
    L.b2m_ingest_host.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_float, C.c_float, vp]
    L.b2m_apply_sform_device.argtypes = [vp, C.POINTER(Result), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.b2m_meshify_raw_host.argtypes = [vp, vp, C.c_int, i64p, C.c_float, C.c_float, C.POINTER(Opts), C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(vp), C.POINTER(vp), C.POINTER(Result)]
    L.b2m_laplacian_hc_host.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    L.b2m_laplacian_hc_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    L.laplacian_smoothHC.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_bool]
    L.laplacian_smoothHC.restype = None
    L.b2m_comm_nccl_id.argtypes = [vp]
    L.b2m_comm_create_nccl.argtypes = [C.POINTER(vp), vp, vp, C.c_int, C.c_int]
    L.b2m_comm_create_local.argtypes = [C.POINTER(vp), C.c_int]
    L.b2m_comm_destroy.argtypes = [vp]
    L.b2m_comm_destroy.restype = None
    L.b2m_comm_reset.argtypes = [vp]
    L.b2m_meshify_slab.argtypes = [vp, vp, vp, i64p, C.c_int64, C.c_int64, C.POINTER(Opts), C.POINTER(SlabResult)]
    L.b2m_meshify_slab_host.argtypes = [vp, vp, vp, i64p, C.c_int64, C.c_int64, C.POINTER(Opts), C.POINTER(vp),
                                        C.POINTER(vp), C.POINTER(SlabResult)]
    L.b2m_stage_smooth.argtypes = [vp, vp, vp, i64p]
    L.b2m_stage_front.argtypes = [vp, vp, i64p, C.POINTER(Opts), vp, vp, C.POINTER(Result)]
    L.b2m_stage_mc.argtypes = [vp, vp, i64p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(Opts), C.POINTER(Result)]
    L.b2m_stage_weld.argtypes = [vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _dims(vol_shape):
    nz, ny, nx = vol_shape
    return (C.c_int64 * 3)(nx, ny, nz)


class DeviceVolume:
    """a float32 volume resident in device memory"""

    def __init__(self, eng, shape, ptr):
        self.eng, self.shape, self.ptr = eng, tuple(shape), ptr
        self.nbytes = int(np.prod(shape)) * 4

    def free(self):
        if self.ptr:
            self.eng.lib.b2m_dev_free(self.ptr)
            self.ptr = None

    def to_host(self):
        out = np.empty(self.shape, np.float32)
        self.eng._chk(self.eng.lib.b2m_d2h(self.eng.ctx, out.ctypes.data, self.ptr, self.nbytes))
        return out


class Engine:
    def __init__(self, device=0):
        self.lib = load()
        self.ctx = C.c_void_p()
        rc = self.lib.b2m_create(C.byref(self.ctx), device)
        if rc != 0:
            raise B2MError(f"b2m_create failed ({rc}): {self.lib.b2m_last_error().decode()}")

    def close(self):
        if self.ctx:
            self.lib.b2m_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc == 1:
            raise MeshifyFailure("meshify failed (reference EXIT_FAILURE semantics)")
        if rc != 0:
            raise B2MError(f"libb2m error {rc}: {self.lib.b2m_last_error().decode()}")

    # ---- measurement ----
    def set_profile(self, on=True):
        self._chk(self.lib.b2m_set_profile(self.ctx, int(on)))

    def kernel_times(self):
        """[(kernel name, ms)] of the last hot-path call (profiling must be on), in launch order"""
        out = []
        for i in range(self.lib.b2m_profile_count(self.ctx)):
            name, ms = C.c_char_p(), C.c_float()
            self._chk(self.lib.b2m_profile_entry(self.ctx, i, C.byref(name), C.byref(ms)))
            out.append((name.value.decode(), float(ms.value)))
        return out

    def timer_start(self):
        self._chk(self.lib.b2m_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._chk(self.lib.b2m_timer_stop(self.ctx, C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._chk(self.lib.b2m_sync(self.ctx))

    def flush_l2(self):
        self._chk(self.lib.b2m_flush_l2(self.ctx))

    # ---- memory ----
    def alloc(self, nbytes):
        p = C.c_void_p()
        self._chk(self.lib.b2m_ctx_alloc(self.ctx, C.byref(p), nbytes))  # on this engine's device, whatever the thread's current one
        return p

    def upload(self, vol):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        p = self.alloc(vol.nbytes)
        self._chk(self.lib.b2m_h2d(self.ctx, p, vol.ctypes.data, vol.nbytes))
        return DeviceVolume(self, vol.shape, p)

    def tiled_volume(self, tile, shape, z_offset=0):
        """periodic replication of a small host tile into a device volume of `shape` (z,y,x), on the device"""
        t = self.upload(tile)
        n = int(np.prod(shape))
        out = DeviceVolume(self, shape, self.alloc(n * 4))
        try:
            self._chk(self.lib.b2m_tile_volume(self.ctx, t.ptr, _dims(t.shape), out.ptr, _dims(shape), int(z_offset)))
        finally:
            t.free()
        return out

    def download(self, ptr, shape, dtype):
        out = np.empty(shape, dtype)
        self._chk(self.lib.b2m_d2h(self.ctx, out.ctypes.data, ptr, out.nbytes))
        return out

    @staticmethod
    def _opts(iso, original_mc, pre_smooth, only_largest, fill_bubbles, backend, verbose=False):
        return Opts(float(iso), int(original_mc), int(pre_smooth), int(only_largest), int(fill_bubbles), int(backend),
                    int(verbose))

    # ---- hot path ----
    def meshify_device(self, dvol, iso, original_mc=0, pre_smooth=True, only_largest=True, fill_bubbles=False,
                       backend=BACKEND_LEWINER, fetch=True, verbose=False):
        """volume already on the device; returns (verts, tris, Result) (verts/tris None if fetch=False)"""
        o = self._opts(iso, original_mc, pre_smooth, only_largest, fill_bubbles, backend, verbose)
        r = Result()
        self._chk(self.lib.b2m_meshify_device(self.ctx, dvol.ptr, _dims(dvol.shape), C.byref(o), C.byref(r)))
        if not fetch:
            return None, None, r
        v = np.empty((r.nverts, 3), np.float64)
        t = np.empty((r.ntris, 3), np.int32)
        self._chk(self.lib.b2m_fetch_mesh(self.ctx, C.byref(r), v.ctypes.data, t.ctypes.data))
        return v, t, r

    def fetch(self, r):
        """(verts, tris) of the device-resident mesh of a Result"""
        v = np.empty((r.nverts, 3), np.float64)
        t = np.empty((r.ntris, 3), np.int32)
        self._chk(self.lib.b2m_fetch_mesh(self.ctx, C.byref(r), v.ctypes.data, t.ctypes.data))
        return v, t

    def meshify(self, vol, iso, original_mc=0, pre_smooth=True, only_largest=True, fill_bubbles=False,
                backend=BACKEND_LEWINER, verbose=False):
        """host volume in, host mesh out (H2D + device pipeline + D2H), like the reference's meshify()."""
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        o = self._opts(iso, original_mc, pre_smooth, only_largest, fill_bubbles, backend, verbose)
        r = Result()
        pv, pt = C.c_void_p(), C.c_void_p()
        self._chk(self.lib.b2m_meshify_host(self.ctx, vol.ctypes.data, _dims(vol.shape), C.byref(o), C.byref(pv),
                                            C.byref(pt), C.byref(r)))
        v = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(r.nverts, 3)).copy()
        t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(r.ntris, 3)).copy()
        _libc.free(pv)
        _libc.free(pt)
        return v, t, r

    NIFTI_DT = {np.dtype(np.uint8): 2, np.dtype(np.int16): 4, np.dtype(np.uint16): 512, np.dtype(np.float32): 16}

    def ingest(self, raw, slope=1.0, inter=0.0):
        """raw voxels (u8 / i16 / u16 / f32 array, z,y,x) -> f32 DeviceVolume, converted on the device like load_nii()"""
        raw = np.ascontiguousarray(raw)
        out = DeviceVolume(self, raw.shape, self.alloc(raw.size * 4))
        self._chk(self.lib.b2m_ingest_host(self.ctx, raw.ctypes.data, self.NIFTI_DT[raw.dtype], raw.size, float(slope),
                                           float(inter), out.ptr))
        return out

    def meshify_raw(self, raw, iso, slope=1.0, inter=0.0, srow=None, original_mc=0, pre_smooth=True, only_largest=True,
                    fill_bubbles=False, backend=BACKEND_LEWINER):
        """raw voxels in, world-space mesh out: ingest + meshify + apply_sform (srow = 3 x 4 floats or None) + D2H"""
        raw = np.ascontiguousarray(raw)
        o = self._opts(iso, original_mc, pre_smooth, only_largest, fill_bubbles, backend)
        r = Result()
        pv, pt = C.c_void_p(), C.c_void_p()
        rows = [None, None, None]
        if srow is not None:
            rows = [(C.c_float * 4)(*[float(x) for x in srow[k]]) for k in range(3)]
        self._chk(self.lib.b2m_meshify_raw_host(self.ctx, raw.ctypes.data, self.NIFTI_DT[raw.dtype], _dims(raw.shape), float(slope),
                                                float(inter), C.byref(o), rows[0], rows[1], rows[2], C.byref(pv), C.byref(pt),
                                                C.byref(r)))
        v = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(r.nverts, 3)).copy()
        t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(r.ntris, 3)).copy()
        _libc.free(pv)
        _libc.free(pt)
        return v, t, r

    # ---- post-smooth ----
    def laplacian_hc(self, verts, tris, iters, alpha=0.1, beta=0.5, lock_edges=True):
        """laplacian_smoothHC() (src/quadric.c:343-394) on a host mesh; returns the smoothed vertices"""
        v = np.ascontiguousarray(verts, dtype=np.float64).copy()
        t = np.ascontiguousarray(tris, dtype=np.int32)
        self._chk(self.lib.b2m_laplacian_hc_host(self.ctx, v.ctypes.data, t.ctypes.data, len(v), len(t), float(alpha), float(beta),
                                                 int(iters), int(lock_edges)))
        return v

    def laplacian_hc_result(self, r, iters, alpha=0.1, beta=0.5, lock_edges=True):
        """the same on the device-resident mesh of a Result (in place, before fetching it)"""
        self._chk(self.lib.b2m_laplacian_hc_device(self.ctx, r.d_verts, r.d_tris, r.nverts, r.ntris, float(alpha), float(beta),
                                                   int(iters), int(lock_edges)))

    def isolevel(self, vol_or_dvol, dark_medium_bright_123):
        """-i d / m / b: the reference's setThreshold() (src/isolevel.c:245-277); 1 = dark, 2 = medium, 3 = bright"""
        iso = C.c_float()
        if isinstance(vol_or_dvol, DeviceVolume):
            n = int(np.prod(vol_or_dvol.shape))
            self._chk(self.lib.b2m_isolevel_device(self.ctx, vol_or_dvol.ptr, n, int(dark_medium_bright_123), C.byref(iso)))
        else:
            v = np.ascontiguousarray(vol_or_dvol, dtype=np.float32)
            self._chk(self.lib.b2m_isolevel_host(self.ctx, v.ctypes.data, v.size, int(dark_medium_bright_123), C.byref(iso)))
        return float(iso.value)

    # ---- atlas front-end (src/nii2mesh.c:492-583) ----
    def atlas_scan(self, dvol):
        """[LabelInfo] for labels 0..nlabel of an indexed volume on the device (one pass)"""
        n, p = C.c_int(), C.POINTER(LabelInfo)()
        self._chk(self.lib.b2m_atlas_scan(self.ctx, dvol.ptr, _dims(dvol.shape), C.byref(n), C.byref(p)))
        out = []
        for i in range(n.value + 1):
            li = LabelInfo()
            C.memmove(C.byref(li), C.byref(p[i]), C.sizeof(LabelInfo))
            out.append(li)
        self.lib.b2m_atlas_free(p)
        return out

    def meshify_label(self, dvol, info, iso=0.5, original_mc=0, pre_smooth=True, fill_bubbles=False,
                      backend=BACKEND_LEWINER, fetch=True):
        """the reference's per-label meshify() (binary volume of the label, -l off) on the label's bounding box"""
        o = self._opts(iso, original_mc, pre_smooth, 0, fill_bubbles, backend)
        r = Result()
        self._chk(self.lib.b2m_meshify_label_device(self.ctx, dvol.ptr, _dims(dvol.shape), C.byref(info), C.byref(o),
                                                    C.byref(r)))
        if not fetch:
            return None, None, r
        v = np.empty((r.nverts, 3), np.float64)
        t = np.empty((r.ntris, 3), np.int32)
        self._chk(self.lib.b2m_fetch_mesh(self.ctx, C.byref(r), v.ctypes.data, t.ctypes.data))
        return v, t, r

    def atlas_meshify_all(self, dvol, iso=0.5, original_mc=0, pre_smooth=True, fill_bubbles=False, backend=BACKEND_LEWINER,
                          workers=8, fetch=True):
        """the whole label loop in ONE call (b2m_atlas_meshify_all): {label: dict(rc, nvox, nverts, ntris, iso_reset,
        verts, tris)} for every label 1..nlabel (rc -100: no voxels, skipped like the reference does)"""
        o = self._opts(iso, original_mc, pre_smooth, 0, fill_bubbles, backend)
        n, p = C.c_int(), C.POINTER(LabelMesh)()
        self._chk(self.lib.b2m_atlas_meshify_all(self.ctx, dvol.ptr, _dims(dvol.shape), C.byref(o), int(workers), int(fetch),
                                                 C.byref(n), C.byref(p)))
        out = {}
        try:
            for i in range(1, n.value + 1):
                m = p[i]
                e = dict(rc=m.rc, nvox=m.nvox, nverts=m.nverts, ntris=m.ntris, iso_reset=m.r.iso_reset, verts=None, tris=None)
                if m.rc == 0 and fetch:
                    e["verts"] = np.ctypeslib.as_array(C.cast(m.verts, C.POINTER(C.c_double)), shape=(m.nverts, 3)).copy()
                    e["tris"] = np.ctypeslib.as_array(C.cast(m.tris, C.POINTER(C.c_int)), shape=(m.ntris, 3)).copy()
                out[i] = e
        finally:
            self.lib.b2m_atlas_meshes_free(p, n.value)
        return out

    def meshify_slab(self, comm, dslab, gshape, z0, iso, original_mc=0, pre_smooth=True, only_largest=True,
                     fill_bubbles=False, backend=BACKEND_LEWINER, verbose=False):
        """this rank's call of the collective z-slab path; dslab = planes [z0, z0+len) of a volume of gshape (z,y,x).
        Returns the SlabResult (blocks stay on the device)."""
        o = self._opts(iso, original_mc, pre_smooth, only_largest, fill_bubbles, backend, verbose)
        r = SlabResult()
        self._chk(self.lib.b2m_meshify_slab(self.ctx, comm, dslab.ptr, _dims(gshape), int(z0), int(dslab.shape[0]),
                                            C.byref(o), C.byref(r)))
        return r

    def nccl_comm(self, id128, rank, world):
        """NCCL transport of the slab path: id128 = bytes from nccl_unique_id() of rank 0"""
        comm = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(id128))
        self._chk(self.lib.b2m_comm_create_nccl(C.byref(comm), self.ctx, buf, rank, world))
        return comm

    def nccl_unique_id(self):
        buf = (C.c_ubyte * 128)()
        self._chk(self.lib.b2m_comm_nccl_id(buf))
        return bytes(buf)

    def meshify_slab_host(self, comm, slab, gshape, z0, iso, original_mc=0, pre_smooth=True, only_largest=True,
                          fill_bubbles=False, backend=BACKEND_LEWINER, hptr=None):
        """host planes in (numpy array, or a raw host pointer `hptr` with slab = its (nz,ny,nx) shape), host blocks out"""
        o = self._opts(iso, original_mc, pre_smooth, only_largest, fill_bubbles, backend)
        r = SlabResult()
        pv, pt = C.c_void_p(), C.c_void_p()
        if hptr is None:
            slab = np.ascontiguousarray(slab, dtype=np.float32)
            hptr, shape = slab.ctypes.data, slab.shape
        else:
            shape = slab
        self._chk(self.lib.b2m_meshify_slab_host(self.ctx, comm, hptr, _dims(gshape), int(z0), int(shape[0]), C.byref(o),
                                                 C.byref(pv), C.byref(pt), C.byref(r)))
        nv = r.nv_edge + r.nv_cent + r.nv_extra
        v = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(max(nv, 1), 3))[:nv].copy()
        t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(max(r.ntris_local, 1), 3))[:r.ntris_local].copy()
        _libc.free(pv)
        _libc.free(pt)
        return r, v, t

    def fetch_slab(self, r):
        """host copies of one rank's blocks: (verts[nv_edge+nv_cent+nv_extra, 3], tris[ntris_local, 3])"""
        nv = r.nv_edge + r.nv_cent + r.nv_extra
        v = self.download(r.d_verts, (nv, 3), np.float64) if nv else np.empty((0, 3), np.float64)
        t = self.download(r.d_tris, (r.ntris_local, 3), np.int32) if r.ntris_local else np.empty((0, 3), np.int32)
        return v, t

    # ---- stage hooks ----
    def smooth(self, vol):
        d_in = self.upload(vol)
        d_out = DeviceVolume(self, d_in.shape, self.alloc(d_in.nbytes))
        rc = self.lib.b2m_stage_smooth(self.ctx, d_in.ptr, d_out.ptr, _dims(d_in.shape))
        if rc not in (0, 1):
            self._chk(rc)
        out = d_out.to_host()
        d_in.free()
        d_out.free()
        return out

    def front(self, vol, iso, pre_smooth=True, only_largest=True, fill_bubbles=False):
        """returns dict(img=composed volume, mask (uint8), iso, lo, hi, mn, mx)"""
        d_in = self.upload(vol)
        n = int(np.prod(d_in.shape))
        d_c = self.alloc(n * 4)
        d_m = self.alloc(n)
        o = self._opts(iso, 0, pre_smooth, only_largest, fill_bubbles, 0)
        r = Result()
        try:
            self._chk(self.lib.b2m_stage_front(self.ctx, d_in.ptr, _dims(d_in.shape), C.byref(o), d_c, d_m, C.byref(r)))
            img = self.download(d_c, d_in.shape, np.float32)
            mask = self.download(d_m, d_in.shape, np.uint8)
        finally:
            d_in.free()
            self.lib.b2m_dev_free(d_c)
            self.lib.b2m_dev_free(d_m)
        return dict(img=img, mask=mask, iso=r.iso_used, lo=list(r.lo), hi=list(r.hi), mn=r.vmin, mx=r.vmax, res=r)

    def mc(self, img, lo, hi, iso, original_mc=0, backend=BACKEND_LEWINER):
        d_in = self.upload(img)
        o = self._opts(iso, original_mc, 0, 0, 0, backend)
        r = Result()
        try:
            self._chk(self.lib.b2m_stage_mc(self.ctx, d_in.ptr, _dims(d_in.shape), (C.c_int * 3)(*lo), (C.c_int * 3)(*hi),
                                            C.byref(o), C.byref(r)))
            v = np.empty((r.nverts, 3), np.float64)
            t = np.empty((r.ntris, 3), np.int32)
            self._chk(self.lib.b2m_fetch_mesh(self.ctx, C.byref(r), v.ctypes.data, t.ctypes.data))
        finally:
            d_in.free()
        return v, t, r

    def weld(self, verts, tris):
        v = np.ascontiguousarray(verts, dtype=np.float64).copy()
        t = np.ascontiguousarray(tris, dtype=np.int32).copy()
        nv, nt = C.c_int(len(v)), C.c_int(len(t))
        self._chk(self.lib.b2m_stage_weld(self.ctx, v.ctypes.data, t.ctypes.data, C.byref(nv), C.byref(nt)))
        return v[:nv.value].copy(), t[:nt.value].copy()


def assemble_slabs(parts):
    """[(SlabResult, verts, tris)] of all ranks -> the assembled (verts, tris) of the whole volume"""
    r0 = parts[0][0].r
    V = np.full((r0.nverts, 3), np.nan, np.float64)
    T = np.full((r0.ntris, 3), -1, np.int32)
    for r, v, t in parts:
        ne, nc, nx = r.nv_edge, r.nv_cent, r.nv_extra
        V[r.v_edge_off:r.v_edge_off + ne] = v[:ne]
        V[r.v_cent_off:r.v_cent_off + nc] = v[ne:ne + nc]
        V[r.v_extra_off:r.v_extra_off + nx] = v[ne + nc:]
        T[r.tri_off:r.tri_off + r.ntris_local] = t
    return V, T


class LocalSlabGroup:
    """`len(cuts)-1` z-slabs of one volume driven by host threads of this process (one b2m_ctx each, all on
    `devices[i]`, default device 0): the single-process transport of the slab path (b2m_comm_create_local)."""

    def __init__(self, world, devices=None):
        self.world = world
        self.engs = [Engine((devices or [0] * world)[i]) for i in range(world)]
        arr = (C.c_void_p * world)()
        self.engs[0]._chk(self.engs[0].lib.b2m_comm_create_local(arr, world))
        self.comms = [C.c_void_p(arr[i]) for i in range(world)]

    def close(self):
        for c in self.comms:
            self.engs[0].lib.b2m_comm_destroy(c)
        self.comms = []
        for e in self.engs:
            e.close()

    def meshify(self, vol, cuts, iso, fetch=True, **flags):
        """vol: host volume (z,y,x); cuts: z boundaries [0, ..., nz].  Returns (verts, tris, [SlabResult])"""
        import threading
        assert len(cuts) == self.world + 1 and cuts[0] == 0 and cuts[-1] == vol.shape[0]
        slabs = [self.engs[i].upload(vol[cuts[i]:cuts[i + 1]]) for i in range(self.world)]
        try:
            return self.meshify_device(slabs, vol.shape, cuts, iso, fetch=fetch, **flags)
        finally:
            for d in slabs:
                d.free()

    def meshify_device(self, slabs, gshape, cuts, iso, fetch=True, **flags):
        import threading
        out, err = [None] * self.world, [None] * self.world

        def work(i):
            try:
                r = self.engs[i].meshify_slab(self.comms[i], slabs[i], gshape, cuts[i], iso, **flags)
                out[i] = (r,) + (self.engs[i].fetch_slab(r) if fetch else (None, None))
            except Exception as ex:  # noqa: BLE001
                err[i] = ex
        th = [threading.Thread(target=work, args=(i,)) for i in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if any(err):
            for c in self.comms:
                self.engs[0].lib.b2m_comm_reset(c)
            raise next(e for e in err if e)
        if not fetch:
            return None, None, [o[0] for o in out]
        V, T = assemble_slabs(out)
        return V, T, [o[0] for o in out]
