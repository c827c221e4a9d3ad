"""Mint the golden vectors under tests/golden/ from the REFERENCE's own code.

The reference ships no tests or golden files for this path, so these fixtures are
outputs of its real templates (src/Influence.h, src/Coefficients.h, src/Kernels.h) compiled from
/root/reference by oracle/Makefile into oracle/_ref/libo3d_ref.so (-O3 -ffp-contract=off), run in the
build container. Run:  python tests/golden/make_golden.py      (needs /root/reference)
The .npz files are committed; the GPU box never needs the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omega3d_b200 import workloads as W  # noqa: E402
from oracle import oracle_py  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def cloud(n, seed, rlo, rhi):
    x, s, _ = W.random_cloud(n, seed=seed)
    return x, s, W.varied_radii(n, seed + 1, rlo, rhi)


def main():
    oracle_py.build(want_ref=True)
    ref = oracle_py.Reference()
    rng = np.random.Generator(np.random.MT19937(99))

    # ---- single-interaction known answers (src/Kernels.h) ----
    s7 = np.array([0.1, 0.2, 0.3, 0.05, 1.0, 0.5, -0.25], f32)
    t4 = np.array([0.4, -0.1, 0.2, 0.05], f32)
    kat = {"s7": s7, "t4": t4, "kernel_0v_0bg": ref.kernel_0v_0bg(s7, t4)}
    tri9 = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], f32)
    str4 = np.array([0, 0, 1, 0], f32)
    for name, t3 in (("near", [0.3, 0.3, 0.2]), ("far", [3.0, -2.0, 5.0]), ("mid", [0.9, 0.9, 0.7]), ("touch", [0.2, 0.1, 0.01])):
        t3 = np.array(t3, f32)
        for g in (0, 1):
            out, fl = ref.rkernel(bool(g), tri9, np.array([0.3, -0.2, 1.0, 0.7], f32) if g else str4, t3, 0.5)
            kat[f"rk_{name}_{'g' if g else 'v'}_t"] = t3
            kat[f"rk_{name}_{'g' if g else 'v'}_out"] = out
            kat[f"rk_{name}_{'g' if g else 'v'}_flops"] = np.int64(fl)
    kat["tri9"], kat["str4_v"], kat["str4_g"] = tri9, str4, np.array([0.3, -0.2, 1.0, 0.7], f32)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)

    # ---- particles -> points, the four kernel variants, non-zero initial outputs (+= semantics) ----
    ns, nt = 601, 257
    sx, ss, sr = cloud(ns, 1001, 0.02, 0.08)
    ss = (ss * f32(ns)).astype(f32)
    tx, _, tr = cloud(nt, 2002, 0.01, 0.06)
    tx[:, :40] = sx[:, :40]  # coincident pairs: zero velocity, non-zero gradient self terms
    u0 = (rng.random((3, nt), dtype=f32) - f32(0.5)).astype(f32)
    g0 = (rng.random((9, nt), dtype=f32) - f32(0.5)).astype(f32)
    pp = {"sx": sx, "ss": ss, "sr": sr, "tx": tx, "tr": tr, "u0": u0, "g0": g0}
    for name, blob, grad in (("0bg", True, True), ("0b", True, False), ("0pg", False, True), ("0p", False, False)):
        tu, tug = u0.copy(), (g0.copy() if grad else None)
        ref.pts_on_pts(sx, sr, ss, tx, tr if blob else None, tu, tug)
        pp["u_" + name] = tu
        if grad:
            pp["g_" + name] = tug
    np.savez_compressed(os.path.join(OUT, "pts_on_pts.npz"), **pp)

    # ---- N=1000 self-influence cloud of SURVEY.md 8c (iii), raw sums (no 1/4pi) ----
    x, s, r = W.random_cloud(1000, seed=12345, radius=0.05)
    tu, tug = np.zeros((3, 1000), f32), np.zeros((9, 1000), f32)
    ref.pts_on_pts(x, r, s, x, r, tu, tug)
    np.savez_compressed(os.path.join(OUT, "self_cloud_1000.npz"), x=x, s=s, r=r, u=tu, g=tug)

    # ---- panels: icosphere with 80 panels, radius 0.5 ----
    nodes_i, idx = W.icosphere(1, 0.5)
    val = W.panel_strengths(idx.shape[0], seed=11, with_source=True)
    area, ts, b1, b2, nrm = ref.surface_props(nodes_i, idx, val)
    nt = 96
    tx = ((rng.random((3, nt), dtype=f32) - f32(0.5)) * f32(2.4)).astype(f32)
    # a third of the targets hug the surface so all three recursion levels are exercised
    for k in range(0, nt, 3):
        v = tx[:, k] / np.linalg.norm(tx[:, k])
        tx[:, k] = (v * (0.5 + 0.004 * (k + 1))).astype(f32)
    u0 = (rng.random((3, nt), dtype=f32) - f32(0.5)).astype(f32)
    g0 = (rng.random((9, nt), dtype=f32) - f32(0.5)).astype(f32)
    pan = {"nodes_i": nodes_i, "idx": idx, "val": val, "area": area, "ts": ts, "b1": b1, "b2": b2, "nrm": nrm,
           "tx": tx, "u0": u0, "g0": g0}
    tu, tug = u0.copy(), g0.copy()
    ref.pan_on_pts(nodes_i, idx, val, tx, None, tu, tug, targ_kind=ref.TARG_FIELD)
    pan["u_grad"], pan["g_grad"] = tu, tug
    tu = u0.copy()
    ref.pan_on_pts(nodes_i, idx, val, tx, None, tu, None, targ_kind=ref.TARG_TRACER)
    pan["u_vel"] = tu
    # particles -> panels (RHS)
    ns = 333
    sx, ss, sr = cloud(ns, 3003, 0.02, 0.05)
    sx = (sx * f32(1.6)).astype(f32)
    sx[:, :60] = (sx[:, :60] / np.linalg.norm(sx[:, :60], axis=0) * f32(0.53)).astype(f32)
    ss = (ss * f32(ns)).astype(f32)
    pu0 = (rng.random((3, idx.shape[0]), dtype=f32) - f32(0.5)).astype(f32)
    pu = pu0.copy()
    ref.pts_on_pan(sx, sr, ss, nodes_i, idx, val, pu)
    pan.update({"psx": sx, "pss": ss, "psr": sr, "pu0": pu0, "pu": pu})
    # panels -> panels through the colocation points
    pan["pan_on_pan_pu"] = ref.pan_on_pan(nodes_i, idx, val, nodes_i, idx, np.zeros_like(val))
    np.savez_compressed(os.path.join(OUT, "panels_80.npz"), **pan)

    # ---- BEM coefficient blocks: 20-panel icosahedron on itself, and on a shifted, smaller copy ----
    n0, i0 = W.icosphere(0, 0.5)
    bc = np.zeros((i0.shape[0], 3), f32)
    a_self = ref.pan_on_pan_coeff(n0, i0, bc)
    n1, i1 = W.icosphere(0, 0.3, center=(0.55, 0.1, -0.05))
    a_cross = ref.pan_on_pan_coeff(n0, i0, bc, target=(n1, i1, bc))
    area0, _, sb1, sb2, snrm = ref.surface_props(n0, i0, bc)
    area1, _, tb1, tb2, tnrm = ref.surface_props(n1, i1, bc)
    np.savez_compressed(os.path.join(OUT, "coeff_20.npz"), n0=n0, i0=i0, n1=n1, i1=i1, a_self=a_self, a_cross=a_cross,
                        area0=area0, sb1=sb1, sb2=sb2, snrm=snrm, area1=area1, tb1=tb1, tb2=tb2, tnrm=tnrm)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
