"""Synthetic inputs shared by the tests and bench.py.

Nothing here is on the hot path: these are the cubes and camera poses the renderer is
measured and checked on, since the reference's own data files are not in its repository.
"""
from __future__ import annotations

import numpy as np

# reference src/main.cpp:37-40
DEFAULT_POS = (5.91813, 6.74116, 17.7067)
DEFAULT_AIM_X = 7.81768

SXI_FOV_H, SXI_FOV_W, SXI_PX = 16.0, 27.0, 0.25  # BASELINE.json configs[0]
SXI_H, SXI_W = 64, 108


def hash01(i, j, k, seed: int) -> np.ndarray:
    """Integer hash of voxel indices -> [0,1), exact in every numpy build (uint32 arithmetic)."""
    with np.errstate(over="ignore"):
        h = (i.astype(np.uint32) * np.uint32(73856093)) ^ (j.astype(np.uint32) * np.uint32(19349663))
        h = h ^ (k.astype(np.uint32) * np.uint32(83492791)) ^ np.uint32((seed * 2654435761) & 0xFFFFFFFF)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return (h >> np.uint32(8)).astype(np.float64) * (1.0 / (1 << 24))


def make_cube(nx: int, ny: int, nz: int, lo=(-40.0, -40.0, -40.0), hi=(40.0, 40.0, 40.0), seed: int = 1234,
              negative_slab: bool = False):
    """Gaussian emissivity shell at 10 RE + floor + white noise, strictly positive.

    Returns (zyx float32 [nz][ny][nx], origin float32[3], spacing float32[3]).
    negative_slab=True makes the voxels with 20 <= x < 22 RE negative, to exercise the
    reference's "first negative sample ends the ray" rule (src/Camera.cpp:207-210).
    """
    n = (nx, ny, nz)
    origin = np.array(lo, dtype=np.float32)
    spacing = np.array([(hi[a] - lo[a]) / (n[a] - 1) for a in range(3)], dtype=np.float32)
    x = origin[0].astype(np.float64) + spacing[0].astype(np.float64) * np.arange(nx)
    y = origin[1].astype(np.float64) + spacing[1].astype(np.float64) * np.arange(ny)
    z = origin[2].astype(np.float64) + spacing[2].astype(np.float64) * np.arange(nz)
    out = np.empty((nz, ny, nx), dtype=np.float32)
    ii = np.arange(nx)[None, :]
    jj = np.arange(ny)[:, None]
    xy2 = (x[None, :] ** 2) + (y[:, None] ** 2)
    for k in range(nz):
        r = np.sqrt(xy2 + z[k] * z[k])
        v = 1e-5 * np.exp(-(((r - 10.0) / 4.0) ** 2)) + 1e-7 + 2e-6 * hash01(ii, jj, np.full((1, 1), k), seed)
        out[k] = v.astype(np.float32)
    if negative_slab:
        sel = (x >= 20.0) & (x < 22.0)
        out[:, :, sel] *= -1.0
    return out, origin, spacing


def config1_cube(negative_slab: bool = False):
    """The reference test cube's shape (tests/DataCube_test.cpp:50-52): 101x161x161, 0.5 RE spacing."""
    return make_cube(101, 161, 161, lo=(-10.0, -40.0, -40.0), hi=(40.0, 40.0, 40.0), negative_slab=negative_slab)


def orbit_poses(n: int, perigee: float = 1.78, apogee: float = 20.0, inclination_deg: float = 73.0) -> np.ndarray:
    """n camera positions, uniform in time, on a SMILE-like Kepler ellipse with its apogee
    over the north pole.  float32 [n][3]; no pose lies on the x axis (y = z = 0)."""
    a = 0.5 * (perigee + apogee)
    e = (apogee - perigee) / (apogee + perigee)
    M = 2.0 * np.pi * (np.arange(n) + 0.5) / n
    E = M.copy()
    for _ in range(60):  # Newton on Kepler's equation
        E -= (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
    xp = a * (np.cos(E) - e)                  # perigee at +xp, apogee at -xp
    yp = a * np.sqrt(1.0 - e * e) * np.sin(E)
    inc = np.deg2rad(inclination_deg)
    u = np.array([np.cos(inc), 0.0, np.sin(inc)])       # apogee direction
    w = np.array([-0.1 * np.sin(inc), 0.99498743710662, 0.1 * np.cos(inc)])  # in-plane, unit, orthogonal to u
    pos = (-xp)[:, None] * u[None, :] + yp[:, None] * w[None, :]
    return pos.astype(np.float32)


def write_dat(path, zyx: np.ndarray, origin, spacing, per_line: int = 10) -> None:
    """Write a cube in the ASCII layout the reference loader parses (src/DataCube.cpp:28-138):
    line 1 meta, line 10 'nx ny nz', then x, y, z coordinates, then nx-value threads."""
    nz, ny, nx = zyx.shape
    coords = [origin[a] + spacing[a] * np.arange(n, dtype=np.float32) for a, n in enumerate((nx, ny, nz))]

    def chunks(vals):
        vals = [repr(float(v)) for v in vals]
        return [" ".join(vals[i:i + per_line]) for i in range(0, len(vals), per_line)]

    with open(path, "w") as f:
        f.write("synthetic emissivity cube (birp_cpp_b200.synthetic)\n")
        for _ in range(8):
            f.write("#\n")
        f.write(f"{nx} {ny} {nz}\n")
        for c in coords:
            for line in chunks(c.astype(np.float32)):
                f.write(line + "\n")
        for k in range(nz):
            for j in range(ny):
                for line in chunks(zyx[k, j]):
                    f.write(line + "\n")
