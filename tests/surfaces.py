"""The ten analytic 60^3 implicit surfaces of the reference's marching-cubes self test
(/root/reference/src/MarchingCubes.c:1282-1362, `compute_data`), restated in numpy with the C
expression's float/double promotion and left-to-right association, so that the volumes are
bit-identical to the ones the compiled self test writes (tools/make_golden.py checks that here).

Known answers (vertices/triangles of Lewiner MC at isolevel 0, recorded from the compiled
reference) are in KNOWN; KNOWN_ORIGINAL is surface 7 with originalMC=1.
"""
import numpy as np

KNOWN = [(302, 600), (792, 1572), (7038, 13720), (554, 1104), (2168, 4352), (5304, 10616), (13196, 25834),
         (11333, 22620), (10800, 20886), (5789, 11298)]
KNOWN_ORIGINAL = {7: (11333, 22732)}
NAMES = ["cushion", "spheres", "plane", "cassini", "blobby", "chair", "cyclide", "2-torus", "mc-case", "drip"]


def surface(kind, n=60):
    f = np.float32
    s = f(n) / f(16)
    t = f(n) / (f(2) * s)
    idx = np.arange(n, dtype=np.float32)
    ax = idx / s - t
    ay = idx / s - (t + f(1.5))
    z, y, x = np.meshgrid(ax, ay, ax, indexing="ij")
    z = np.ascontiguousarray(z); y = np.ascontiguousarray(y); x = np.ascontiguousarray(x)
    d = np.float64
    if kind == 0:
        v = z * z * x * x - z * z * z * z - 2 * z * x * x + 2 * z * z * z + x * x - z * z - (x * x - z) * (x * x - z) \
            - y * y * y * y - 2 * x * x * y * y - y * y * z * z + 2 * y * y * z + y * y
    elif kind == 1:
        v = ((x - 2) * (x - 2) + (y - 2) * (y - 2) + (z - 2) * (z - 2) - 1) \
            * ((x + 2) * (x + 2) + (y - 2) * (y - 2) + (z - 2) * (z - 2) - 1) \
            * ((x - 2) * (x - 2) + (y + 2) * (y + 2) + (z - 2) * (z - 2) - 1)
    elif kind == 2:
        v = x + y + z - 3
    elif kind == 3:
        q = f(0.45) * f(0.45)
        v = (x * x + y * y + z * z + q) * (x * x + y * y + z * z + q) - f(16) * f(0.45) * f(0.45) * (x * x + z * z) \
            - f(0.5) * f(0.5)
    elif kind == 4:
        a = x * x * x * x - 5 * x * x + y * y * y * y - 5 * y * y + z * z * z * z - 5 * z * z
        v = (a.astype(d) + 11.8).astype(f)
    elif kind == 5:
        k = f(0.95) * f(25)
        v = (x * x + y * y + z * z - k) * (x * x + y * y + z * z - k) \
            - f(0.8) * ((z - 5) * (z - 5) - 2 * x * x) * ((z + 5) * (z + 5) - 2 * y * y)
    elif kind == 6:
        b, dd, a, c = f(2), f(6), f(2), f(3)
        p = x * x + y * y + z * z + b * b - dd * dd
        v = p * p - 4 * ((a * x - c * dd) * (a * x - c * dd) + b * b * y * y)
    elif kind == 7:
        R, r = f(4), f(1.85)
        p = x * x + y * y + z * z + R * R - r * r
        q = x * x + (y + R) * (y + R) + z * z + R * R - r * r
        v = (p * p - 4 * R * R * (x * x + y * y)) * (q * q - 4 * R * R * ((y + R) * (y + R) + z * z))
    elif kind == 8:
        mx, my, mz = (1 - x).astype(d), (1 - y).astype(d), (1 - z).astype(d)
        X, Y, Z = x.astype(d), y.astype(d), z.astype(d)
        v = -26.5298 * mx * my * mz + 81.9199 * X * my * mz - 100.68 * X * Y * mz + 3.5498 * mx * Y * mz \
            + 24.1201 * mx * my * Z - 74.4702 * X * my * Z + 91.5298 * X * Y * Z - 3.22998 * mx * Y * Z
        v = v.astype(f)
    elif kind == 9:
        A = (x * x + y * y).astype(d)
        Z = z.astype(d)
        inner = 0.995 * Z * Z + 0.005 - (z * z * z).astype(d)
        v = (A - 0.5 * inner + 0.0025).astype(f)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(v.astype(f))
