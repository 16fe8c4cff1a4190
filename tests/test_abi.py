"""CPU tests of the drop-in boundary: libb2m.so builds for sm_100a, loads without a GPU, exports
every symbol include/*.h declares, fails loudly (no CPU fallback) when no device is present, and
the host-side meshify.h utilities behave like the reference's (src/meshify.c:973-982, :1021-1045)."""
import ctypes as C
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    syms = []
    for h in ("b2m.h", "meshify.h", "isolevel.h", "quadric.h"):
        txt = (ROOT / "include" / h).read_text()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms += re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", txt)
    return sorted(set(syms))


def test_header_symbols_exported(libb2m):
    syms = declared_symbols()
    assert "meshify" in syms and "b2m_meshify_device" in syms and "laplacian_smoothHC" in syms and len(syms) >= 25
    for s in syms:
        assert hasattr(libb2m, s), f"libb2m.so does not export {s}"


def test_only_sm100a_code(libb2m):
    from nii2mesh_b200 import lib
    out = subprocess.run(["cuobjdump", "-lelf", str(lib.LIBPATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device(libb2m):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert libb2m.b2m_create(C.byref(ctx), 0) != 0
    assert b"no CPU fallback" in libb2m.b2m_last_error()
    from nii2mesh_b200 import lib
    with pytest.raises(lib.B2MError):
        lib.Engine(0)
    # the reference-facing meshify() returns EXIT_FAILURE instead of computing on the CPU
    img = np.random.default_rng(0).standard_normal((8, 8, 8)).astype(np.float32)
    dim = (C.c_short * 3)(8, 8, 8)
    pt, pp, nt, nv = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    libb2m.meshify.argtypes = [C.c_void_p, C.POINTER(C.c_short), C.c_int, C.c_float, C.POINTER(C.c_void_p),
                               C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_bool, C.c_bool,
                               C.c_bool, C.c_bool]
    rc = libb2m.meshify(img.ctypes.data, dim, 0, 0.0, C.byref(pt), C.byref(pp), C.byref(nt), C.byref(nv), True, True,
                        False, False)
    assert rc == 1 and not pt.value and not pp.value


def test_product_never_imports_oracle():
    for p in (ROOT / "nii2mesh_b200").rglob("*"):
        if p.suffix in (".py", ".c", ".cu", ".cuh", ".h"):
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, p


def _sform_args(libobj):
    libobj.apply_sform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_float)]
    libobj.apply_sform.restype = None
    libobj.strip_ext.argtypes = [C.c_char_p]
    libobj.strip_ext.restype = None


def _apply(libobj, v, t, sx, sy, sz):
    v, t = v.copy(), t.copy()
    f4 = C.c_float * 4
    libobj.apply_sform(t.ctypes.data, v.ctypes.data, len(t), len(v), f4(*sx), f4(*sy), f4(*sz))
    return v, t


def test_apply_sform_and_strip_ext(libb2m):
    _sform_args(libb2m)
    rng = np.random.default_rng(5)
    v = rng.uniform(0, 200, (500, 3))
    t = rng.integers(0, 500, (900, 3)).astype(np.int32)
    rows = [([-1.25, 0.01, 0, 90.0], [0.02, 1.25, 0, -126.0], [0, 0.03, 1.25, -72.0]),   # negative determinant proxy
            ([1.0, 0, 0, -10.0], [0, 1.0, 0, 5.0], [0, 0, 1.0, 2.5])]
    ref = None
    try:
        from oracle import Ref, ref_available
        if ref_available():
            ref = Ref("lewiner").lib
            _sform_args(ref)
    except Exception:  # noqa: BLE001
        ref = None
    for sx, sy, sz in rows:
        gv, gt = _apply(libb2m, v, t, sx, sy, sz)
        m = np.array([sx, sy, sz], np.float32).astype(np.float64)
        want = v @ m[:, :3].T + m[:, 3]
        assert np.allclose(gv, want, rtol=1e-12, atol=1e-9)
        flip = np.prod(m[:, :3].sum(axis=1)) < 0
        assert np.array_equal(gt, t[:, [1, 0, 2]] if flip else t)
        if ref is not None:
            rv, rt = _apply(ref, v, t, sx, sy, sz)
            assert np.array_equal(gv, rv) and np.array_equal(gt, rt)
    for name in ("a/b.nii", "a/b.nii.gz", "noext", "dir.d/file", "./x.mz3", ".hidden"):
        b1 = C.create_string_buffer(name.encode(), 64)
        libb2m.strip_ext(b1)
        if ref is not None:
            b2 = C.create_string_buffer(name.encode(), 64)
            ref.strip_ext(b2)
            assert b1.value == b2.value, name
    b = C.create_string_buffer(b"a/b.nii.gz", 64)
    libb2m.strip_ext(b)
    assert b.value == b"a/b.nii"


def test_host_copy_kernels(libb2m):
    """hostcopy.c (the pool threads' non-temporal copy / f32->f64 widening of the D2H leg): every alignment and tail"""
    libb2m.b2m_stream_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    libb2m.b2m_stream_copy.restype = None
    libb2m.b2m_stream_widen.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    libb2m.b2m_stream_widen.restype = None
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, 70000, dtype=np.uint8)
    for n in (0, 1, 31, 4095, 4096, 4097, 65536, 69999):
        for so in (0, 1, 7):
            for do in (0, 3, 32):
                if so + n > len(src):
                    continue
                dst = np.full(n + 128, 0xAB, np.uint8)
                libb2m.b2m_stream_copy(dst.ctypes.data + do, src.ctypes.data + so, n)
                assert np.array_equal(dst[do:do + n], src[so:so + n]) and (dst[:do] == 0xAB).all() and (dst[do + n:] == 0xAB).all()
    f = rng.standard_normal(5000).astype(np.float32)
    f[:8] = [0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, 1.17549435e-38]
    for n in (0, 1, 3, 4, 15, 16, 17, 1000, 4999):
        for do in (0, 1, 3):
            dst = np.full(n + 8, 7.0, np.float64)
            libb2m.b2m_stream_widen(dst.ctypes.data + 8 * do, f.ctypes.data + 4, n)
            assert np.array_equal(dst[do:do + n].view(np.uint64), f[1:1 + n].astype(np.float64).view(np.uint64))
            assert (dst[:do] == 7.0).all() and (dst[do + n:] == 7.0).all()


@pytest.mark.timeout(120)
def test_copy_pool_concurrent_callers_do_not_deadlock(libb2m):
    """a pre-fault job used to keep the process-wide pool locked until its owner returned from the collective
    pipeline, so a second rank of a single-process slab group blocked in its own copy and never reached the barrier.
    Several callers, blocks >= 64 MiB (the pre-fault threshold), a barrier between pre-fault and copy; no GPU needed."""
    libb2m.b2m_pool_selftest.argtypes = [C.c_size_t, C.c_int]
    libb2m.b2m_set_copy_threads.argtypes = [C.c_int]
    assert libb2m.b2m_set_copy_threads(4) in (0, 1)
    assert libb2m.b2m_get_copy_threads() >= 1
    for callers in (1, 2, 3):
        assert libb2m.b2m_pool_selftest(40 << 20, callers) == 0, callers
