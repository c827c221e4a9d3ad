"""Seeded synthetic inputs for the Biot-Savart path.

All arrays are float32 SoA in the layout of the reference's element containers
(``Points``: x[3], s[3], r - src/Points.h:54-140; ``Surfaces``: node x[3], idx, area, ts[3] -
src/Surfaces.h:62-225). Nothing here computes an influence; it only makes inputs.
"""
from __future__ import annotations

import numpy as np

CLOUD_SEED = 20240517


def random_cloud(n: int, seed: int | None = None, radius: float | None = None):
    """Uniform vortex-particle cloud in [-0.5,0.5]^3: strengths U[-0.5,0.5]^3 / n, radius 1.5 n^(-1/3).

    Returns x (3,n), s (3,n), r (n,) float32, C-contiguous rows.
    """
    rng = np.random.Generator(np.random.MT19937(CLOUD_SEED + n if seed is None else seed))
    x = rng.random((3, n), dtype=np.float32) - np.float32(0.5)
    s = (rng.random((3, n), dtype=np.float32) - np.float32(0.5)) / np.float32(n)
    r = np.full(n, (1.5 * n ** (-1.0 / 3.0)) if radius is None else radius, np.float32)
    return np.ascontiguousarray(x), np.ascontiguousarray(s.astype(np.float32)), r


def varied_radii(n: int, seed: int, lo: float, hi: float):
    rng = np.random.Generator(np.random.MT19937(seed))
    return (lo + (hi - lo) * rng.random(n, dtype=np.float32)).astype(np.float32)


def vortex_rings(n: int, seed: int = 7, separation: float = 0.25, major: float = 0.35, minor: float = 0.06):
    """Two coaxial thick-cored rings (the leapfrog / colliding-ring geometry of
    3Dexamples/leapfrog_vortex_rings_nv.json), n particles in total, vorticity along the ring tangent."""
    rng = np.random.Generator(np.random.MT19937(seed))
    half = n // 2
    xs, ss = [], []
    for k, cnt in enumerate((half, n - half)):
        theta = rng.random(cnt) * 2.0 * np.pi
        rho = minor * np.sqrt(rng.random(cnt))
        phi = rng.random(cnt) * 2.0 * np.pi
        rad = major + rho * np.cos(phi)
        z = (k - 0.5) * separation + rho * np.sin(phi)
        xs.append(np.stack([rad * np.cos(theta), rad * np.sin(theta), z]))
        circ = 1.0 / cnt
        ss.append(circ * 2.0 * np.pi * major * np.stack([-np.sin(theta), np.cos(theta), np.zeros(cnt)]))
    x = np.concatenate(xs, axis=1).astype(np.float32)
    s = np.concatenate(ss, axis=1).astype(np.float32)
    r = np.full(n, 1.5 * (2.0 * np.pi * major * np.pi * minor ** 2 * 2 / n) ** (1.0 / 3.0), np.float32)
    return np.ascontiguousarray(x), np.ascontiguousarray(s), r


def icosphere(levels: int, radius: float = 0.5, center=(0.0, 0.0, 0.0)):
    """Triangulated sphere by midpoint refinement of an icosahedron (20 * 4^levels panels), outward
    normals. Returns nodes (nn,3) float32 interleaved and idx (np,3) uint32 - the ElementPacket layout
    (src/ElementPacket.h:31-37) the reference's Surfaces ctor consumes."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(levels):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    nodes = (np.array(verts) * radius + np.array(center)).astype(np.float32)
    return np.ascontiguousarray(nodes), np.ascontiguousarray(np.array(faces, np.uint32))


def panel_strengths(npan: int, seed: int = 11, with_source: bool = True):
    """Per-panel (vortex-sheet x1, vortex-sheet x2, source-sheet) strengths, interleaved (np,3)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    v = (rng.random((npan, 3), dtype=np.float32) - np.float32(0.5)).astype(np.float32)
    if not with_source:
        v[:, 2] = 0.0
    return np.ascontiguousarray(v)


def strided_subset(n: int, count: int):
    """`count` evenly strided target indices out of n (the CPU-baseline / parity sample)."""
    count = min(n, count)
    return (np.arange(count, dtype=np.int64) * n) // count
