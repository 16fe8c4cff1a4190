#!/usr/bin/env python3
"""end-to-end meshify() timing (pinned host volume in, malloc'd mesh out) for host-path experiments:
    [B2M_RING_CHUNK_KB=.. B2M_RING_SLOTS=.. B2M_COPY_THREADS=..] python tools/e2e_probe.py [size] [reps]"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = lib.Engine(0)
L = eng.lib
hp = C.c_void_p()
eng._chk(L.b2m_host_alloc(C.byref(hp), n ** 3 * 4))
hvol = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(n, n, n))
t = synth.gyroid_tile(128)
hvol.reshape(n // 128, 128, n // 128, 128, n // 128, 128)[...] = t[None, :, None, :, None, :]
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]
o = lib.Opts(0.0, 0, 1, 1, 1, 0, 0)
rows = []
for rep in range(reps + 2):
    r = lib.Result()
    pv, pt = C.c_void_p(), C.c_void_p()
    t0 = time.perf_counter()
    eng._chk(L.b2m_meshify_host(eng.ctx, hp, (C.c_int64 * 3)(n, n, n), C.byref(o), C.byref(pv), C.byref(pt), C.byref(r)))
    ms = (time.perf_counter() - t0) * 1e3
    libc.free(pv)
    libc.free(pt)
    if rep >= 2:
        rows.append((ms, r.h2d_ms, r.ms[7], r.d2h_ms))
a = np.array(rows)
print({k: os.environ.get(k) for k in ("B2M_RING_CHUNK_KB", "B2M_RING_SLOTS", "B2M_COPY_THREADS", "B2M_H2D_OVERLAP")},
      "total %.1f h2d %.1f device %.1f d2h %.1f (ms, mean of %d)" % (*a.mean(axis=0), len(a)))
