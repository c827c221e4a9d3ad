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


# ---- the example cases' initial conditions (3Dexamples/*.json), restated from src/FlowFeature.cpp and src/Simulation.cpp ----
def sim_scales(re: float, dt: float, overlap_ratio: float = 1.5, core_size_ratio: float = 8.0 ** 0.5):
    """Simulation::get_hnu / get_ips / get_vdelta (src/Simulation.cpp:57-58,77-79): nominal particle spacing and core radius."""
    f = np.float32
    hnu = f(np.sqrt(f(f(dt) / f(re))))
    ips = f(f(core_size_ratio) * hnu)
    vdelta = f(f(overlap_ratio) * ips)
    return float(ips), float(vdelta)


def _onb(normal):
    """normalizeVec + branchlessONB (src/MathHelper.h:186-192,220-230), float32 as the reference."""
    f = np.float32
    n = np.asarray(normal, f)
    n = (n * f(1.0 / np.sqrt(np.float64(f(f(n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]))))).astype(f)
    sign = f(np.copysign(1.0, n[2]))
    a = f(-1.0 / np.float64(sign + n[2]))
    b = f(f(n[0] * n[1]) * a)
    b1 = np.array([f(1.0 + np.float64(f(f(sign * n[0]) * n[0]) * a)), f(sign * b), f(-sign * n[0])], f)
    b2 = np.array([b, f(sign + f(f(n[1] * n[1]) * a)), f(-n[1])], f)
    return n, b1, b2


def singular_ring(center=(0.0, 0.0, 0.0), normal=(1.0, 0.0, 0.0), majrad=0.5, circ=1.0, ips=0.015):
    """SingularRing::init_elements (src/FlowFeature.cpp:789-832): one row of particles on a circle, strength along
    the tangent. Returns x, s (3,n) float32 (agrees with the reference's generator to float rounding; the fixtures
    in tests/golden/convection.npz hold the reference's own output)."""
    f = np.float32
    ndiam = int(1 + (2.0 * np.pi * f(majrad)) / f(ips))
    this_ips = f((2.0 * np.pi * f(majrad)) / f(ndiam))
    _, b1, b2 = _onb(normal)
    theta = (2.0 * np.pi * np.arange(ndiam, dtype=f).astype(np.float64) / float(ndiam)).astype(f)
    ct, st = np.cos(theta).astype(f), np.sin(theta).astype(f)
    c = np.asarray(center, f)
    x = np.stack([c[d] + f(majrad) * (b1[d] * ct + b2[d] * st) for d in range(3)]).astype(f)
    s = np.stack([f(this_ips * f(circ)) * (b2[d] * ct - b1[d] * st) for d in range(3)]).astype(f)
    return np.ascontiguousarray(x), np.ascontiguousarray(s)


def thick_ring(center=(0.0, 0.0, 0.0), normal=(1.0, 0.0, 0.0), majrad=0.5, minrad=0.05, circ=1.0, ips=0.015):
    """ThickRing::init_elements (src/FlowFeature.cpp:941-1019): a disk of concentric particle layers swept around the ring."""
    f = np.float32
    nlayers = int(1 + f(minrad) / f(ips))
    dx, dy, dl = [0.0], [0.0], [1.0]
    for l in range(1, nlayers):
        rad = f(l) * f(ips)
        nl = int(1 + (2.0 * np.pi * rad) / f(ips))
        phi = (2.0 * np.pi * np.arange(nl, dtype=f).astype(np.float64) / float(nl)).astype(f)
        dx += list(rad * np.cos(phi).astype(f)); dy += list(rad * np.sin(phi).astype(f))
        dl += list((f(majrad) + rad * np.cos(phi).astype(f)) / f(majrad))
    dx, dy, dl = np.asarray(dx, f), np.asarray(dy, f), np.asarray(dl, f)
    ndisk = dx.size
    ndiam = int(1 + (2.0 * np.pi * f(majrad)) / f(ips))
    this_ips = f((2.0 * np.pi * f(majrad)) / f(ndiam))
    nrm, b1, b2 = _onb(normal)
    theta = (2.0 * np.pi * np.arange(ndiam, dtype=f).astype(np.float64) / float(ndiam)).astype(f)
    ct, st = np.cos(theta).astype(f)[:, None], np.sin(theta).astype(f)[:, None]
    c = np.asarray(center, f)
    x = np.stack([(c[d] + (f(majrad) + dx[None, :]) * (b1[d] * ct + b2[d] * st) + dy[None, :] * nrm[d]).reshape(-1) for d in range(3)]).astype(f)
    sscale = (dl * this_ips * f(circ) / f(ndisk)).astype(f)[None, :]
    s = np.stack([(sscale * (b2[d] * ct - b1[d] * st)).reshape(-1) for d in range(3)]).astype(f)
    return np.ascontiguousarray(x), np.ascontiguousarray(s)


EXAMPLES = {
    # 3Dexamples/single_vortex_ring_nv.json (BASELINE configs[0])
    "single_vortex_ring_nv": dict(re=71.11111450195313, dt=0.0020000000949949026, fs=(0.0, 0.0, 0.0), rings=[
        dict(center=(0.0, 0.0, 0.0), normal=(0.8999999761581421, 0.05000000074505806, 0.10000000149011612), majrad=0.5, circ=1.0)]),
    # 3Dexamples/leapfrog_vortex_rings_nv.json (BASELINE configs[1] before growing the particle count)
    "leapfrog_vortex_rings_nv": dict(re=40.000003814697266, dt=0.0020000000949949026, fs=(0.0, 0.0, 0.0), rings=[
        dict(center=(0.0, 0.0, 0.0), normal=(0.8999999761581421, 0.05000000074505806, 0.10000000149011612), majrad=0.5, circ=1.0),
        dict(center=(-0.18000000715255737, -0.009999999776482582, -0.019999999552965164),
             normal=(0.8999999761581421, 0.05000000074505806, 0.10000000149011612), majrad=0.5, circ=1.0)]),
    # 3Dexamples/colliding_vortex_rings_nv.json (BASELINE configs[2], its inviscid form: "viscous": "none")
    "colliding_vortex_rings_nv": dict(re=40.000003814697266, dt=0.0020000000949949026, fs=(0.0, 0.0, 0.0), rings=[
        dict(center=(0.4000000059604645, 0.0, 0.0), normal=(-0.8999999761581421, -0.05000000074505806, -0.10000000149011612),
             majrad=0.5, circ=1.0),
        dict(center=(0.03999999910593033, -0.019999999552965164, -0.03999999910593033),
             normal=(0.8999999761581421, 0.05000000074505806, 0.10000000149011612), majrad=0.5, circ=1.0)]),
}


def example_case(name: str, minrad: float | None = None, ips: float | None = None):
    """Initial particles of an example input file as Simulation::add_elements builds them (src/Simulation.cpp:1009):
    every flow structure's particles appended to one collection, radius = vdelta. `minrad` swaps the singular rings
    for thick ones and `ips` overrides the spacing - the way the BASELINE configs grow the particle count.
    Returns x, s (3,n), r (n,), dt, fs."""
    case = EXAMPLES[name]
    ips0, vdelta = sim_scales(case["re"], case["dt"])
    if ips is not None:
        vdelta = float(np.float32(1.5) * np.float32(ips))
    else:
        ips = ips0
    xs, ss = [], []
    for ring in case["rings"]:
        x, s = (singular_ring(ips=ips, **ring) if minrad is None else thick_ring(ips=ips, minrad=minrad, **ring))
        xs.append(x); ss.append(s)
    x, s = np.concatenate(xs, axis=1), np.concatenate(ss, axis=1)
    return np.ascontiguousarray(x), np.ascontiguousarray(s), np.full(x.shape[1], vdelta, np.float32), case["dt"], case["fs"]


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
