"""Deterministic synthetic volumes and a minimal NIfTI-1 reader.

Arrays are float32 indexed [z, y, x] (x fastest), i.e. the reference's `v = x + y*NX + z*NX*NY`
layout (/root/reference/src/meshify.c:186-189).
"""
import gzip
import struct

import numpy as np


def noisy_sphere(n, seed=1234, noise=2.0):
    """S<n> "noisy sphere" (BASELINE config 2 at n=512): float32 arithmetic throughout, isolevel 0."""
    z, y, x = np.meshgrid(*(np.arange(n, dtype=np.float32),) * 3, indexing="ij", sparse=True)
    c = np.float32((n - 1) / 2)
    rng = np.random.Generator(np.random.Philox(seed))
    r = np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    v = (np.float32(0.35 * n) - r) + np.float32(noise) * rng.standard_normal((n, n, n), dtype=np.float32)
    return np.ascontiguousarray(v.astype(np.float32))


def gyroid_tile(P=128):
    """One P^3 period of the "gyroid + bumps" field, computed in float64 and cast once to float32."""
    i = np.arange(P, dtype=np.float64)
    z, y, x = np.meshgrid(i, i, i, indexing="ij", sparse=True)
    k = 2 * np.pi / P
    g = np.sin(k * x) * np.cos(k * y) + np.sin(k * y) * np.cos(k * z) + np.sin(k * z) * np.cos(k * x)

    def d2(u, c):
        return np.minimum(np.abs(u - c), P - np.abs(u - c)) ** 2

    def bump(c):
        return np.exp(-(d2(x, c) + d2(y, c) + d2(z, c)) / 50.0)

    return (g - 3.0 * bump(16.0) + 3.0 * bump(48.0)).astype(np.float32)


def gyroid(n, P=128):
    """G<n> "gyroid + bumps" (config 3 at n=1024, config 5 at n=2048): tile replicated, isolevel 0.
    n need not be a multiple of P (the tiling is cropped)."""
    tile = gyroid_tile(P)
    reps = -(-n // P)
    v = np.tile(tile, (reps, reps, reps))[:n, :n, :n]
    return np.ascontiguousarray(v)


def random_blobs(shape, seed=0, smooth=2, thresh=0.0):
    """Small random multi-component test volume: smoothed white noise (many clusters + bubbles)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(shape).astype(np.float32)
    for _ in range(smooth):
        for ax in range(3):
            v = (np.roll(v, 1, ax) + v + np.roll(v, -1, ax)) / np.float32(3)
    v = v / np.float32(v.std())
    return np.ascontiguousarray((v - np.float32(thresh)).astype(np.float32))


def load_nifti(path):
    """Minimal NIfTI-1 reader mirroring load_nii (/root/reference/src/nii2mesh.c:53-175):
    u8/i16/u16/f32, native endian, optional gzip, `raw*scl_slope + scl_inter` in float32.
    Returns (volume[z,y,x] float32, header dict with dim, pixdim, srow_x/y/z)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if struct.unpack_from("<i", raw, 0)[0] != 348:
        raise ValueError("only native-endian NIfTI-1 is supported")
    dim = struct.unpack_from("<8h", raw, 40)
    datatype = struct.unpack_from("<h", raw, 70)[0]
    pixdim = struct.unpack_from("<8f", raw, 76)
    vox_offset = struct.unpack_from("<f", raw, 108)[0]
    slope, inter = struct.unpack_from("<2f", raw, 112)
    srow = struct.unpack_from("<12f", raw, 280)
    dt = {2: np.uint8, 4: np.int16, 512: np.uint16, 16: np.float32}.get(datatype)
    if dt is None:
        raise ValueError(f"unsupported datatype {datatype}")
    nx, ny, nz = dim[1], dim[2], dim[3]
    n = nx * ny * nz
    off = int(round(vox_offset))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off)
    if slope == 0.0:
        slope = 1.0
    vol = (data.astype(np.float32) * np.float32(slope)) + np.float32(inter)
    hdr = dict(dim=(nx, ny, nz), pixdim=pixdim, srow_x=srow[0:4], srow_y=srow[4:8], srow_z=srow[8:12],
               datatype=datatype, scl_slope=slope, scl_inter=inter)
    return np.ascontiguousarray(vol.reshape(nz, ny, nx).astype(np.float32)), hdr
