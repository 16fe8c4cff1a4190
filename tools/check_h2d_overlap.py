#!/usr/bin/env python3
"""EXPERIMENTAL path, not yet run on a GPU: B2M_H2D_OVERLAP=1 sends a pinned volume up in z-chunks on a second stream
and lets the smooth follow the transfer (pipeline.cu: b2m_meshify_host, b2m_front_run).  This script is the check to run
before turning it on: same mesh as the plain path, bit for bit, and the end-to-end times of both.
    B2M_H2D_OVERLAP=1 python tools/check_h2d_overlap.py [size]      (the flag is read once per process)"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = lib.Engine(0)
L = eng.lib
hp = C.c_void_p()
eng._chk(L.b2m_host_alloc(C.byref(hp), n ** 3 * 4))
hvol = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(n, n, n))
hvol[...] = synth.gyroid(n) if n <= 256 else np.tile(synth.gyroid_tile(128), (n // 128,) * 3)
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]
o = lib.Opts(0.0, 0, 1, 1, 1, 0, 0)
out = []
for rep in range(4):
    r = lib.Result()
    pv, pt = C.c_void_p(), C.c_void_p()
    t0 = time.perf_counter()
    eng._chk(L.b2m_meshify_host(eng.ctx, hp, (C.c_int64 * 3)(n, n, n), C.byref(o), C.byref(pv), C.byref(pt), C.byref(r)))
    ms = (time.perf_counter() - t0) * 1e3
    v = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(r.nverts, 3)).copy()
    t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int)), shape=(r.ntris, 3)).copy()
    libc.free(pv)
    libc.free(pt)
    out.append((ms, r.h2d_ms, r.ms[7], r.d2h_ms))
print("overlap" if os.environ.get("B2M_H2D_OVERLAP") else "plain", n, "last call: total %.1f ms, h2d %.1f, device %.1f, d2h %.1f" % out[-1])
# reference result through the device-resident path (no host copies involved)
d = eng.upload(hvol)
v2, t2, _ = eng.meshify_device(d, 0.0, 0, 1, 1, 1, 0)
assert np.array_equal(t, t2) and np.array_equal(v.view(np.uint64), v2.view(np.uint64)), "MISMATCH"
print("identical to the device-resident path:", len(v), "vertices", len(t), "triangles")
