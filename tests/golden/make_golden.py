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


def convection(ref):
    """tests/golden/convection.npz: the reference's own Points::finalize_vels / move and the Convection::advect
    sequence (oracle/ref_driver.cpp: o3d_ref_finalize_vels, o3d_ref_move, o3d_ref_advect), plus the initial
    conditions of the two particle-only example inputs from the reference's SingularRing::init_elements."""
    rng = np.random.Generator(np.random.MT19937(4711))
    g = {}
    # --- finalize_vels and the three move() overloads on random stage data ---
    n = 257
    x = (rng.random((3, n), dtype=f32) - f32(0.5)).astype(f32)
    s = ((rng.random((3, n), dtype=f32) - f32(0.5)) * f32(0.01)).astype(f32)
    s[:, 5] = 0.0   # a zero-strength particle: the elongation update is skipped (circmagsqrd == 0)
    elong = (f32(1.0) + rng.random(n, dtype=f32) * f32(0.2)).astype(f32)
    us = [(rng.random((3, n), dtype=f32) - f32(0.5)).astype(f32) for _ in range(3)]
    gs = [((rng.random((9, n), dtype=f32) - f32(0.5)) * f32(4.0)).astype(f32) for _ in range(3)]
    g.update(mv_x=x, mv_s=s, mv_elong=elong, mv_dt=np.float64(0.0125), mv_fs=np.array([0.3, -0.1, 0.05]))
    for k in range(3):
        g[f"mv_u{k}"], g[f"mv_g{k}"] = us[k], gs[k]
    fu, fg = us[0].copy(), gs[0].copy()
    ref.finalize_vels(fu, fg, g["mv_fs"])
    g["fin_u"], g["fin_g"] = fu, fg
    for order, wt in ((1, [0.75]), (2, [0.25, 0.75]), (3, [2.0 / 9.0, 3.0 / 9.0, 4.0 / 9.0])):
        a, b, c = x.copy(), s.copy(), elong.copy()
        uo = np.zeros((3, n), f32)
        ref.move(order, float(g["mv_dt"]), wt, us[:order], gs[:order], a, b, c, uo)
        g[f"mv{order}_wt"] = np.array(wt)
        g[f"mv{order}_x"], g[f"mv{order}_s"], g[f"mv{order}_elong"], g[f"mv{order}_u"] = a, b, c, uo
    # no gradients on one stage: advection only, strengths and elongation untouched
    a, b, c = x.copy(), s.copy(), elong.copy()
    ref.move(2, float(g["mv_dt"]), [0.5, 0.5], us[:2], [gs[0], None], a, b, c, None)
    g["mv2ng_x"], g["mv2ng_s"], g["mv2ng_elong"] = a, b, c
    # --- Convection::advect, all three orders, random cloud with a freestream ---
    n = 400
    cx, cs, cr = W.random_cloud(n, seed=777, radius=0.08)
    cs = (cs * f32(40.0)).astype(f32)
    g.update(adv_x=cx, adv_s=cs, adv_r=cr, adv_dt=np.float64(0.05), adv_fs=np.array([0.1, 0.0, 0.2]), adv_steps=np.int64(2))
    for order in (1, 2, 3):
        a, b, e = cx.copy(), cs.copy(), np.ones(n, f32)
        u, ug = ref.advect(order, 2, 0.05, g["adv_fs"], a, b, cr, e)
        g[f"adv{order}_x"], g[f"adv{order}_s"], g[f"adv{order}_elong"], g[f"adv{order}_u"], g[f"adv{order}_ug"] = a, b, e, u, ug
    g["adv_stats"] = np.array(ref.stats(b, e), f32)
    # --- the particle-only example inputs: initial conditions from the reference's generator, then 5 RK2 steps ---
    for name in ("single_vortex_ring_nv", "leapfrog_vortex_rings_nv"):
        case = W.EXAMPLES[name]
        ips, vdelta = W.sim_scales(case["re"], case["dt"])
        parts = [ref.singular_ring(r["center"], r["normal"], r["majrad"], r["circ"], ips) for r in case["rings"]]
        x0 = np.ascontiguousarray(np.concatenate([p[0] for p in parts], axis=1))
        s0 = np.ascontiguousarray(np.concatenate([p[1] for p in parts], axis=1))
        r0 = np.full(x0.shape[1], vdelta, f32)
        a, b, e = x0.copy(), s0.copy(), np.ones(x0.shape[1], f32)
        u, ug = ref.advect(2, 5, case["dt"], case["fs"], a, b, r0, e)
        g.update({f"{name}_x0": x0, f"{name}_s0": s0, f"{name}_r0": r0, f"{name}_x": a, f"{name}_s": b, f"{name}_elong": e,
                  f"{name}_u": u, f"{name}_ug": ug, f"{name}_steps": np.int64(5)})
    # a thick ring (the geometry the BASELINE configs grow to 1M-4M particles), initial condition only
    tx, ts = ref.thick_ring((0.1, 0.0, 0.0), (0.9, 0.05, 0.1), 0.5, 0.07, 1.0, 0.03)
    g["thick_x0"], g["thick_s0"] = tx, ts
    np.savez_compressed(os.path.join(OUT, "convection.npz"), **g)


def reflect(ref):
    """tests/golden/reflect.npz: the reference's reflect_panp2 / clear_inner_panp2 (src/Reflect.h) on its real
    Surfaces<float> / Points<float>: the 320-panel sphere of flow_over_sphere.json and a particle cloud around and
    inside it, with particles placed under nodes and edge midpoints so that the tie branch of the hit list runs."""
    rng = np.random.Generator(np.random.MT19937(2718))
    nodes_i, idx = W.icosphere(2, 0.5)
    val = np.zeros((idx.shape[0], 3), f32)
    area, ts, b1, b2, nrm = ref.surface_props(nodes_i, idx, val)
    nodes = np.ascontiguousarray(nodes_i.T)
    nt = 3000
    x = ((rng.random((3, nt), dtype=f32) - f32(0.5)) * f32(1.4)).astype(f32)
    x[:, :60] = (nodes[:, :60] * f32(0.98)).astype(f32)                                             # under nodes: 5-6 panels tie
    x[:, 60:120] = ((nodes[:, idx[:60, 0]] + nodes[:, idx[:60, 1]]) * f32(0.495)).astype(f32)       # under edge midpoints: 2 tie
    x[:, 120:180] = (nodes[:, 100:160] * f32(1.01)).astype(f32)                                     # just outside
    g = {"nodes_i": nodes_i, "idx": idx, "nrm": nrm, "x0": x}
    a = x.copy()
    g["reflect_moved"] = np.int64(ref.reflect(nodes_i, idx, a))
    g["reflect_x"] = a
    cm, ips = f32(0.5 / np.sqrt(2.0 * np.pi)), f32(0.0894427)        # the call of src/Convection.h:260 with the sphere case's ips
    a = x.copy()
    g["clear_moved"] = np.int64(ref.clear_inner(1, nodes_i, idx, a, np.full(nt, 0.03, f32), cm, ips))
    g["clear_x"], g["clear_cm"], g["clear_ips"] = a, cm, ips
    a = g["reflect_x"].copy()                                         # Diffusion.h:292-306: reflect, then clear
    g["both_moved"] = np.int64(ref.clear_inner(1, nodes_i, idx, a, np.full(nt, 0.03, f32), f32(0.2), f32(0.05)))
    g["both_x"] = a
    np.savez_compressed(os.path.join(OUT, "reflect.npz"), **g)


def vtk(ref):
    """tests/golden/vtk.npz: the bytes of the file the reference's Points<float>::write_vtk writes (src/Points.h:851-1039)."""
    import tempfile
    rng = np.random.Generator(np.random.MT19937(31))
    g = {}
    for n in (1, 37):
        x = (rng.random((3, n), dtype=f32) - f32(0.5)).astype(f32)
        s = (rng.random((3, n), dtype=f32) - f32(0.5)).astype(f32)
        r = (f32(0.01) + rng.random(n, dtype=f32) * f32(0.1)).astype(f32)
        u = (rng.random((3, n), dtype=f32) - f32(0.5)).astype(f32)
        with tempfile.TemporaryDirectory() as d:
            data = ref.write_vtk(x, s, r, u, 3, 42, 0.0625 * n, d)
        g.update({f"x{n}": x, f"s{n}": s, f"r{n}": r, f"u{n}": u, f"time{n}": np.float64(0.0625 * n),
                  f"file{n}": np.frombuffer(data, np.uint8)})
    np.savez_compressed(os.path.join(OUT, "vtk.npz"), **g)


def status(ref):
    """tests/golden/status.npz: ElementBase::get_total_circ / Points::get_total_impulse of the reference's Points<float>
    (oracle/ref_driver.cpp: o3d_ref_totals) on three collections, and the bytes of the files the reference's StatusFile writes
    when driven like Simulation::dump_stats_to_status (o3d_ref_status_lines): two data sets in one file, both formats."""
    import tempfile
    g = {}
    cases = {"ring": W.example_case("single_vortex_ring_nv")[:2],
             "leap": W.example_case("leapfrog_vortex_rings_nv", minrad=0.05, ips=0.015)[:2],
             "cloud": W.random_cloud(20000, seed=77)[:2]}
    for name, (x, s) in cases.items():
        if name == "cloud":
            s = (s * f32(1000.0)).astype(f32)
        c, i = ref.totals(x, s)
        g.update({f"{name}_x": x, f"{name}_s": s, f"{name}_circ": c, f"{name}_imp": i})
    # the totals AFTER the reference's own convection steps (Convection::advect order 2 through its Points methods,
    # o3d_ref_advect): 10 steps of the single ring as shipped, 100 steps of the thick leapfrogging rings. (The thin 210-particle
    # ring is not a conservation case: after 100 steps the reference's own two builds differ by 0.07 in total circulation;
    # the thick rings conserve circulation to 5e-9 of sum|s| and impulse to 1.3e-3.)
    for name, steps in (("ring", 10), ("leap", 100)):
        x0, s0, r, dt, fs = (W.example_case("single_vortex_ring_nv") if name == "ring" else
                             W.example_case("leapfrog_vortex_rings_nv", minrad=0.05, ips=0.015))
        x, s, e = x0.copy(), s0.copy(), np.ones(x0.shape[1], f32)
        ref.advect(2, steps, dt, fs, x, s, r, e)
        c, i = ref.totals(x, s)
        g.update({f"{name}_steps": np.int32(steps), f"{name}_dt": np.float64(dt), f"{name}_r": r, f"{name}_circ_after": c, f"{name}_imp_after": i})
    rng = np.random.Generator(np.random.MT19937(5))
    nl = 7
    vals = (rng.standard_normal((nl, 7)) * np.array([1, 1e-6, 1e-3, 1, 10, 1e4, 1e-9])).astype(f32)
    vals[:, 0] = (np.arange(nl) % 4 * 0.002).astype(f32)     # time restarts with the second data set
    vals[3, 1:] = [0.0, -0.0, 1.0, 123456.0, 1234567.0, 1e-5]  # the %g corner cases
    nv = np.array([210, 210, 215, 1048524, 0, 7, 123456789], np.int32)
    reset = np.array([0, 0, 0, 0, 1, 0, 0], np.int32)
    g.update(status_vals=vals, status_nv=nv, status_reset=reset)
    for fmt, tag in ((0, "dat"), (1, "csv")):
        with tempfile.TemporaryDirectory() as d:
            data = ref.status_lines(os.path.join(d, "status." + tag), fmt, vals, nv, reset)
        g["status_" + tag] = np.frombuffer(data, np.uint8)
    np.savez_compressed(os.path.join(OUT, "status.npz"), **g)


def cores():
    """tests/golden/cores.npz: particles -> points from the three builds of the reference whose src/CoreFunc.h has
    another core function #defined (oracle/Makefile: libo3d_ref_{rm,exp,v2}.so) - the four kernel variants each,
    per-particle radii, coincident pairs, += on non-zero outputs - plus a uniform-radius self-influence cloud."""
    rng = np.random.Generator(np.random.MT19937(9090))
    ns, nt = 601, 257
    sx, ss, sr = cloud(ns, 1001, 0.02, 0.08)
    ss = (ss * f32(ns)).astype(f32)
    tx, _, tr = cloud(nt, 2002, 0.01, 0.06)
    tx[:, :40] = sx[:, :40]  # coincident pairs
    # a ladder of close pairs (separation / radius from 0.02 to 3): walks the exponential core through all three
    # of its branches (reld3 < 0.001, in between, > 16: src/CoreFunc.h:114-128)
    for k in range(40, 100):
        tx[:, k] = sx[:, k] + f32(0.02 + 0.05 * (k - 40)) * sr[k] * np.array([0.6, -0.48, 0.64], f32)
    u0 = (rng.random((3, nt), dtype=f32) - f32(0.5)).astype(f32)
    g0 = (rng.random((9, nt), dtype=f32) - f32(0.5)).astype(f32)
    g = {"sx": sx, "ss": ss, "sr": sr, "tx": tx, "tr": tr, "u0": u0, "g0": g0}
    x, s, r = W.random_cloud(1000, seed=12345, radius=0.05)
    g.update(cx=x, cs=s, cr=r)
    for core, tag in ((1, "rm"), (2, "exp"), (3, "v2")):
        ref = oracle_py.Reference(core=core)
        for name, blob, grad in (("0bg", True, True), ("0b", True, False), ("0pg", False, True), ("0p", False, False)):
            tu, tug = u0.copy(), (g0.copy() if grad else None)
            ref.pts_on_pts(sx, sr, ss, tx, tr if blob else None, tu, tug)
            g[f"{tag}_u_{name}"] = tu
            if grad:
                g[f"{tag}_g_{name}"] = tug
        tu, tug = np.zeros((3, 1000), f32), np.zeros((9, 1000), f32)
        ref.pts_on_pts(x, r, s, x, r, tu, tug)
        g[f"{tag}_cloud_u"], g[f"{tag}_cloud_g"] = tu, tug
        # panels -> points (with gradients) on the inputs of panels_80.npz: the panel leaves evaluate the core at zero
        # radius, where the four cores coincide up to rounding
        pg = np.load(os.path.join(OUT, "panels_80.npz"))
        tu, tug = pg["u0"].copy(), pg["g0"].copy()
        ref.pan_on_pts(pg["nodes_i"], pg["idx"], pg["val"], pg["tx"], None, tu, tug, targ_kind=ref.TARG_FIELD)
        g[f"{tag}_pan_u"], g[f"{tag}_pan_g"] = tu, tug
        # two Convection::advect steps of each order through that build's own Points methods (o3d_ref_advect)
        for order in (1, 2, 3):
            ax, as_, ar = W.random_cloud(300, seed=11, radius=0.08)
            as_ = (as_ * f32(30.0)).astype(f32)
            ae = np.ones(300, f32)
            if core == 1 and order == 1:
                g["adv_x0"], g["adv_s0"], g["adv_r"] = ax.copy(), as_.copy(), ar
            ref.advect(order, 2, 0.02, (0.1, 0.0, 0.0), ax, as_, ar, ae)
            g[f"{tag}_adv{order}_x"], g[f"{tag}_adv{order}_s"], g[f"{tag}_adv{order}_elong"] = ax, as_, ae
    np.savez_compressed(os.path.join(OUT, "cores.npz"), **g)


def main():
    oracle_py.build(want_ref=True)
    ref = oracle_py.Reference()
    if len(sys.argv) > 1 and sys.argv[1] == "vtk":
        vtk(ref)
        print("vtk.npz", os.path.getsize(os.path.join(OUT, "vtk.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "reflect":
        reflect(ref)
        print("reflect.npz", os.path.getsize(os.path.join(OUT, "reflect.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "status":
        status(ref)
        print("status.npz", os.path.getsize(os.path.join(OUT, "status.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "convection":
        convection(ref)
        print("convection.npz", os.path.getsize(os.path.join(OUT, "convection.npz")))
        return
    convection(ref)
    reflect(ref)
    vtk(ref)
    status(ref)
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
    if sys.argv[1:] == ["cores"]:
        cores()
    else:
        main()
        cores()
