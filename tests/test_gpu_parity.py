"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Tolerances are BASELINE.json's: max |err| / max |ref| <= 1e-5 on velocity, <= 1e-4 on the velocity gradient
(conftest.VEL_TOL / GRAD_TOL). Integer work (the reference's flop bookkeeping = its leaf/split counts) is exact.
"""
import numpy as np
import pytest

from conftest import GRAD_TOL, VEL_TOL, golden, rel_err

from omega3d_b200 import influence as I
from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32


def soa(nodes_i):
    return np.ascontiguousarray(nodes_i.T)


# ---- golden vectors (outputs of the reference's own code) -------------------------------------------------
@pytest.mark.parametrize("variant,blob,grad", [("0bg", True, True), ("0b", True, False), ("0pg", False, True), ("0p", False, False)])
def test_pts_on_pts_golden(cuda_ctx, variant, blob, grad):
    g = golden("pts_on_pts.npz")
    tu = g["u0"].copy()
    tug = g["g0"].copy() if grad else None
    cuda_ctx.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], g["tr"] if blob else None, tu, tug)
    assert rel_err(tu, g["u_" + variant]) <= VEL_TOL
    if grad:
        assert rel_err(tug, g["g_" + variant]) <= GRAD_TOL
    ns, nt = g["sx"].shape[1], g["tx"].shape[1]
    per = {"0bg": 70, "0b": 33, "0pg": 68, "0p": 31}[variant]
    assert cuda_ctx.flops == nt * ((12 if grad else 3) + per * ns)  # src/Influence.h:310,366,475,534


def test_kat_single_interaction(cuda_ctx):
    k = golden("kat.npz")
    s7, t4 = k["s7"], k["t4"]
    sx = s7[:3].reshape(3, 1).copy(); sr = s7[3:4].copy(); ss = s7[4:7].reshape(3, 1).copy()
    tx = t4[:3].reshape(3, 1).copy(); tr = t4[3:4].copy()
    tu, tug = np.zeros((3, 1), f32), np.zeros((9, 1), f32)
    cuda_ctx.pts_on_pts(sx, sr, ss, tx, tr, tu, tug)
    ref = k["kernel_0v_0bg"]
    np.testing.assert_allclose(tu[:, 0], ref[:3], rtol=2e-6)
    np.testing.assert_allclose(tug[:, 0], ref[3:], rtol=4e-6)


def test_self_cloud_golden(cuda_ctx):
    g = golden("self_cloud_1000.npz")
    p = I.Points(g["x"], g["s"], g["r"], I.active, I.lagrangian)
    p.zero_vels()
    I.points_affect_points(p, p, I.ResultsType(I.velandgrad), I.ExecEnv(), cuda_ctx)  # source aliases target
    assert rel_err(p.u, g["u"]) <= VEL_TOL and rel_err(p.ug, g["g"]) <= GRAD_TOL


def test_panels_golden(cuda_ctx, restate):
    g = golden("panels_80.npz")
    surf = I.Surfaces(soa(g["nodes_i"]), g["idx"], g["val"], I.active, I.fixed)
    fld = I.Points(g["tx"], e=I.inert, m=I.fixed)
    fld.u[:], fld.ug[:] = g["u0"], g["g0"]
    I.panels_affect_points(surf, fld, I.ResultsType(I.velonly), I.ExecEnv(), cuda_ctx)
    assert rel_err(fld.u - g["u0"], g["u_grad"] - g["u0"]) <= VEL_TOL
    assert rel_err(fld.ug - g["g0"], g["g_grad"] - g["g0"]) <= GRAD_TOL
    # leaf-for-leaf identical traversal: the reference's flop count is an exact function of it
    flops = 12 * g["tx"].shape[1]
    nodes = soa(g["nodes_i"])
    for i in range(g["tx"].shape[1]):
        for j in range(g["idx"].shape[0]):
            tri9 = g["nodes_i"][g["idx"][j]].reshape(9)
            str4 = np.array([surf.ts[0, j] / surf.area[j], surf.ts[1, j] / surf.area[j], surf.ts[2, j] / surf.area[j], g["val"][j, 2]], f32)
            flops += restate.rkernel(True, tri9, str4, g["tx"][:, i], float(surf.area[j]))[1]
    assert cuda_ctx.flops == flops

    trc = I.Points(g["tx"], e=I.inert, m=I.lagrangian)
    trc.u[:] = g["u0"]
    I.panels_affect_points(surf, trc, I.ResultsType(I.velonly), I.ExecEnv(), cuda_ctx)
    assert rel_err(trc.u - g["u0"], g["u_vel"] - g["u0"]) <= VEL_TOL

    src = I.Points(g["psx"], g["pss"], g["psr"], I.active, I.lagrangian)
    surf.pu[:] = g["pu0"]
    I.points_affect_panels(src, surf, I.ResultsType(I.velonly), I.ExecEnv(), cuda_ctx)
    assert rel_err(surf.pu - g["pu0"], g["pu"] - g["pu0"]) <= VEL_TOL

    tsurf = I.Surfaces(soa(g["nodes_i"]), g["idx"], np.zeros_like(g["val"]), I.reactive, I.fixed)
    tsurf.nrm = g["nrm"]  # the reference ctor's normals: the colocation points sit 1e-4 off the sheet, 1 ulp matters
    I.panels_affect_panels(surf, tsurf, I.ResultsType(I.velonly), I.ExecEnv(), cuda_ctx)
    assert rel_err(tsurf.pu, g["pan_on_pan_pu"]) <= VEL_TOL


def test_panel_work_pool_modes_agree(cuda_ctx):
    """o3d_cuda_set_panel_queue: 0 = every lane walks its own deferred pairs, 1 (default) = panels -> points pools them per warp,
    2 = particles -> panels as well. Same leaves and splits (the reference's flop count is an exact function of the traversal),
    same sums up to the regrouping of FP32 terms inside a tile."""
    g = golden("panels_80.npz")
    surf = I.Surfaces(soa(g["nodes_i"]), g["idx"], g["val"], I.active, I.fixed)
    src = I.Points(g["psx"], g["pss"], g["psr"], I.active, I.lagrangian)
    out = {}
    try:
        for mode in (0, 1, 2):
            cuda_ctx.set_panel_queue(mode)
            fld = I.Points(g["tx"], e=I.inert, m=I.fixed)
            I.panels_affect_points(surf, fld, I.ResultsType(I.velonly), I.ExecEnv(), cuda_ctx)
            f1 = cuda_ctx.flops
            surf.pu[:] = 0
            I.points_affect_panels(src, surf, I.ResultsType(I.velonly), I.ExecEnv(), cuda_ctx)
            out[mode] = (fld.u.copy(), fld.ug.copy(), surf.pu.copy(), f1, cuda_ctx.flops)
    finally:
        cuda_ctx.set_panel_queue(1)
    for mode in (1, 2):
        assert out[mode][3] == out[0][3] and out[mode][4] == out[0][4]
        assert rel_err(out[mode][0], out[0][0]) <= 2e-6 and rel_err(out[mode][1], out[0][1]) <= 2e-6
        assert rel_err(out[mode][2], out[0][2]) <= 2e-6
    assert np.array_equal(out[1][2], out[0][2])      # mode 1 leaves particles -> panels on the per-lane kernel
    assert rel_err(out[0][2], g["pu"] - g["pu0"]) <= VEL_TOL


def test_coeff_golden(cuda_ctx):
    g = golden("coeff_20.npz")
    s0 = I.Surfaces(soa(g["n0"]), g["i0"], None, I.reactive, I.fixed)
    s1 = I.Surfaces(soa(g["n1"]), g["i1"], None, I.reactive, I.fixed)
    # bases from the reference ctor so the comparison isolates the kernel
    s0.b1, s0.b2, s0.nrm, s0.area = g["sb1"], g["sb2"], g["snrm"], g["area0"]
    s1.b1, s1.b2, s1.nrm, s1.area = g["tb1"], g["tb2"], g["tnrm"], g["area1"]
    a = I.panels_on_panels_coeff(s0, s0, cuda_ctx)
    assert a.shape == g["a_self"].shape and rel_err(a, g["a_self"]) <= 2e-5
    d = a.reshape(60, 60, order="F")
    fac = 1.0 / (4.0 * np.pi)
    for j in range(20):  # the diagonal override, src/Coefficients.h:414-436
        blk = d[3 * j:3 * j + 3, 3 * j:3 * j + 3]
        np.testing.assert_allclose(blk, np.array([[0, -2 * np.pi, 0], [2 * np.pi, 0, 0], [0, 0, 2 * np.pi]]) * fac, rtol=1e-6)
    a = I.panels_on_panels_coeff(s0, s1, cuda_ctx)
    assert rel_err(a, g["a_cross"]) <= 2e-5


# ---- seeded inputs against the oracle -----------------------------------------------------------------------
@pytest.mark.parametrize("ns,nt", [(1, 1), (2, 3), (511, 513), (512, 256), (4099, 2050), (30000, 777)])
@pytest.mark.parametrize("blob,grad", [(True, True), (True, False), (False, True), (False, False)])
def test_pts_on_pts_vs_oracle(cuda_ctx, restate, ns, nt, blob, grad):
    sx, ss, _ = W.random_cloud(ns, seed=100 + ns)
    sr = W.varied_radii(ns, 200 + ns, 0.5 * ns ** (-1 / 3), 2.0 * ns ** (-1 / 3))
    tx, _, _ = W.random_cloud(nt, seed=300 + nt)
    tr = W.varied_radii(nt, 400 + nt, 0.5 * ns ** (-1 / 3), 2.0 * ns ** (-1 / 3)) if blob else None
    a_u, b_u = np.zeros((3, nt), f32), np.zeros((3, nt), f32)
    a_g, b_g = (np.zeros((9, nt), f32), np.zeros((9, nt), f32)) if grad else (None, None)
    cuda_ctx.pts_on_pts(sx, sr, ss, tx, tr, a_u, a_g)
    restate.pts_on_pts(sx, sr, ss, tx, tr, b_u, b_g)
    assert rel_err(a_u, b_u) <= VEL_TOL
    if grad:
        assert rel_err(a_g, b_g) <= GRAD_TOL


@pytest.mark.parametrize("ns,nt,grad", [(5003, 250001, True), (777, 130000, True), (1500, 400003, False), (513, 114000, True)])
def test_whole_blocks_then_tail_shapes_vs_oracle(cuda_ctx, restate, ns, nt, grad):
    """Target counts above one sweep of the persistent CTAs (148 x 768 with gradients, 148 x 1536 without): every CTA
    finishes whole target blocks in place (phase A), then takes its share of the stream-K tail, with odd tile counts and a
    ragged last block; the initial outputs are non-zero (+=). A strided sample of the targets against the oracle, and every
    target's trace-free gradient."""
    sx, ss, _ = W.random_cloud(ns, seed=100 + ns)
    sr = W.varied_radii(ns, 200 + ns, 0.5 * ns ** (-1 / 3), 2.0 * ns ** (-1 / 3))
    tx, _, _ = W.random_cloud(nt, seed=300 + nt)
    tr = W.varied_radii(nt, 400 + nt, 0.5 * ns ** (-1 / 3), 2.0 * ns ** (-1 / 3))
    rng = np.random.Generator(np.random.MT19937(nt))
    u0 = (rng.standard_normal((3, nt)) * 1e-3).astype(f32)
    g0 = (rng.standard_normal((9, nt)) * 1e-2).astype(f32) if grad else None
    a_u, a_g = u0.copy(), (g0.copy() if grad else None)
    cuda_ctx.pts_on_pts(sx, sr, ss, tx, tr, a_u, a_g)
    sel = np.unique(np.concatenate([W.strided_subset(nt, 400), np.arange(nt - 40, nt), np.arange(0, 40)]))
    b_u = np.ascontiguousarray(u0[:, sel])
    b_g = np.ascontiguousarray(g0[:, sel]) if grad else None
    restate.pts_on_pts(sx, sr, ss, np.ascontiguousarray(tx[:, sel]), np.ascontiguousarray(tr[sel]), b_u, b_g)
    assert rel_err(a_u[:, sel] - u0[:, sel], b_u - u0[:, sel]) <= VEL_TOL
    if grad:
        assert rel_err(a_g[:, sel] - g0[:, sel], b_g - g0[:, sel]) <= GRAD_TOL
        d = a_g - g0
        assert np.max(np.abs(d[0] + d[4] + d[8])) <= 2e-4 * np.max(np.abs(d))
    assert np.all(np.isfinite(a_u))


def test_empty_inputs_are_noops(cuda_ctx):
    x0 = np.zeros((3, 0), f32); r0 = np.zeros(0, f32)
    x, s, r = W.random_cloud(10, seed=5)
    u = np.ones((3, 10), f32); g = np.ones((9, 10), f32)
    cuda_ctx.pts_on_pts(x0, r0, x0, x, r, u, g)   # no sources
    assert np.all(u == 1) and np.all(g == 1)
    cuda_ctx.pts_on_pts(x, r, s, x0, r0, np.zeros((3, 0), f32), None)  # no targets
    nodes, idx = W.icosphere(0)
    surf = I.Surfaces(soa(nodes), idx, W.panel_strengths(20), I.active)
    cuda_ctx.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, None, x0, np.zeros((3, 0), f32), None)
    pu = np.ones((3, 20), f32)
    cuda_ctx.pts_on_pan(x0, x0, surf.x, surf.idx, surf.area, pu)
    assert np.all(pu == 1)


def test_bad_arguments_return_errors(cuda_ctx):
    x, s, r = W.random_cloud(10, seed=5)
    u = np.zeros((3, 10), f32)
    lib = cuda_ctx.lib
    rc = lib.o3d_cuda_pts_on_pts(cuda_ctx.h, 10, None, None, None, None, None, None, None, 10, None, None, None, None,
                                 None, None, None, None, None)
    assert rc == 1 and b"NULL" in lib.o3d_cuda_last_error(cuda_ctx.h)
    nodes, idx = W.icosphere(0)
    bad = idx.copy(); bad[3, 1] = 999
    with pytest.raises(I.O3DError):
        cuda_ctx.pan_on_pts(soa(nodes), bad, np.zeros((3, 20), f32), np.ones(20, f32), None, x, u, None)
    with pytest.raises(I.O3DError):  # the reference asserts on this combination (src/Influence.h:368-370)
        I.points_affect_points(I.Points(x, s, r), I.Points(x, e=I.inert, m=I.lagrangian), I.ResultsType(I.velandgrad), I.ExecEnv(), cuda_ctx)


@pytest.mark.parametrize("levels,nt,grad", [(1, 500, True), (2, 3000, False), (2, 1500, True)])
def test_pan_on_pts_vs_oracle(cuda_ctx, restate, levels, nt, grad):
    nodes, idx = W.icosphere(levels, 0.5)
    surf = I.Surfaces(soa(nodes), idx, W.panel_strengths(idx.shape[0], seed=21), I.active)
    rng = np.random.Generator(np.random.MT19937(77))
    tx = ((rng.random((3, nt), dtype=f32) - f32(0.5)) * f32(1.6)).astype(f32)
    tx[:, ::2] = (tx[:, ::2] / np.linalg.norm(tx[:, ::2], axis=0) * (0.5 + 0.2 * rng.random(tx[:, ::2].shape[1]) ** 3)).astype(f32)
    a_u, b_u = np.zeros((3, nt), f32), np.zeros((3, nt), f32)
    a_g, b_g = (np.zeros((9, nt), f32), np.zeros((9, nt), f32)) if grad else (None, None)
    cuda_ctx.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], tx, a_u, a_g)
    restate.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], tx, b_u, b_g)
    assert rel_err(a_u, b_u) <= VEL_TOL
    if grad:
        assert rel_err(a_g, b_g) <= GRAD_TOL


@pytest.mark.parametrize("levels,ns", [(1, 1000), (2, 20000)])
def test_pts_on_pan_vs_oracle(cuda_ctx, restate, levels, ns):
    nodes, idx = W.icosphere(levels, 0.5)
    surf = I.Surfaces(soa(nodes), idx, None, I.reactive)
    sx, ss, _ = W.random_cloud(ns, seed=31)
    sx = (sx * f32(1.5)).astype(f32)
    sx[:, ::3] = (sx[:, ::3] / np.linalg.norm(sx[:, ::3], axis=0) * f32(0.52)).astype(f32)
    a, b = np.zeros((3, surf.np_), f32), np.zeros((3, surf.np_), f32)
    cuda_ctx.pts_on_pan(sx, ss, surf.x, surf.idx, surf.area, a)
    restate.pts_on_pan(sx, ss, surf.x, surf.idx, surf.area, b)
    assert rel_err(a, b) <= VEL_TOL


def test_coeff_vs_oracle_sphere_320(cuda_ctx, restate):
    """The flow_over_sphere body (C4): 320 panels, 960 x 960 block."""
    nodes, idx = W.icosphere(2, 0.5)
    s = I.Surfaces(soa(nodes), idx, None, I.reactive)
    a = I.panels_on_panels_coeff(s, s, cuda_ctx)
    b = restate.pan_on_pan_coeff(s.x, s.idx, s.b1, s.b2, s.area, s.x, s.idx, s.b1, s.b2, s.nrm, s.area, True)
    assert rel_err(a, b) <= 2e-5


# ---- size-independent properties at BASELINE.json's sizes ---------------------------------------------------
@pytest.fixture(scope="module")
def big():
    n = 1 << 20   # configs[1]: ~1M particles
    x, s, r = W.random_cloud(n)
    return n, x, s, r


def test_1m_subsample_vs_oracle_and_sharding(cuda_ctx, restate, big):
    n, x, s, r = big
    u, g = np.zeros((3, n), f32), np.zeros((9, n), f32)
    cuda_ctx.pts_on_pts(x, r, s, x, r, u, g)
    t = cuda_ctx.last_timing()
    assert t["launches"] >= 2 and t["kernel_ms"] > 0
    sel = W.strided_subset(n, 256)
    tx = np.ascontiguousarray(x[:, sel]); tr = np.ascontiguousarray(r[sel])
    ru, rg = np.zeros((3, sel.size), f32), np.zeros((9, sel.size), f32)
    restate.pts_on_pts(x, r, s, tx, tr, ru, rg)
    assert rel_err(u[:, sel], ru) <= VEL_TOL and rel_err(g[:, sel], rg) <= GRAD_TOL
    # trace-free gradient (vortex-only sources)
    assert np.max(np.abs(g[0] + g[4] + g[8])) <= 1e-4 * np.max(np.abs(g))
    # target sharding is exact: any target slice evaluated alone reproduces the full run bit for bit
    lo, hi = n // 2 - 1000, n // 2 + 1000
    su, sg = np.zeros((3, hi - lo), f32), np.zeros((9, hi - lo), f32)
    cuda_ctx.pts_on_pts(x, r, s, np.ascontiguousarray(x[:, lo:hi]), np.ascontiguousarray(r[lo:hi]), su, sg)
    assert rel_err(su, u[:, lo:hi]) <= 2e-7 and rel_err(sg, g[:, lo:hi]) <= 2e-7


def test_linearity_and_accumulation(cuda_ctx):
    n = 50000
    x, s, r = W.random_cloud(n)
    tx = np.ascontiguousarray(x[:, :4096]); tr = np.ascontiguousarray(r[:4096])
    u1, g1 = np.zeros((3, 4096), f32), np.zeros((9, 4096), f32)
    cuda_ctx.pts_on_pts(x, r, s, tx, tr, u1, g1)
    u2, g2 = np.zeros((3, 4096), f32), np.zeros((9, 4096), f32)
    cuda_ctx.pts_on_pts(x, r, (s * f32(4)).astype(f32), tx, tr, u2, g2)
    assert np.array_equal(u2, u1 * f32(4)) and np.array_equal(g2, g1 * f32(4))  # power-of-two scaling is exact
    # two source collections accumulate like one (the `+=` contract of find_vels, src/Convection.h:144)
    h = n // 2
    ua, ga = np.zeros((3, 4096), f32), np.zeros((9, 4096), f32)
    cuda_ctx.pts_on_pts(np.ascontiguousarray(x[:, :h]), r[:h].copy(), np.ascontiguousarray(s[:, :h]), tx, tr, ua, ga)
    cuda_ctx.pts_on_pts(np.ascontiguousarray(x[:, h:]), r[h:].copy(), np.ascontiguousarray(s[:, h:]), tx, tr, ua, ga)
    assert rel_err(ua, u1) <= 1e-6 and rel_err(ga, g1) <= 1e-6


def test_translation_invariance_and_self_term(cuda_ctx):
    n = 20000
    x, s, r = W.random_cloud(n, seed=9)
    u1, g1 = np.zeros((3, n), f32), np.zeros((9, n), f32)
    cuda_ctx.pts_on_pts(x, r, s, x, r, u1, g1)
    xs = (x + np.array([[0.25], [-0.5], [0.125]], f32)).astype(f32)  # exact in float for these magnitudes? no: compare loosely
    u2, g2 = np.zeros((3, n), f32), np.zeros((9, n), f32)
    cuda_ctx.pts_on_pts(xs, r, s, xs, r, u2, g2)
    assert rel_err(u2, u1) <= 1e-5 and rel_err(g2, g1) <= 1e-4
    # one particle on itself: zero velocity, non-zero antisymmetric gradient (src/Kernels.h:184-192)
    p = np.array([[0.1], [0.2], [0.3]], f32); w = np.array([[1.0], [0.5], [-0.25]], f32); rad = np.array([0.05], f32)
    u, g = np.zeros((3, 1), f32), np.zeros((9, 1), f32)
    cuda_ctx.pts_on_pts(p, rad, w, p, rad, u, g)
    assert np.all(u == 0)
    assert g[0, 0] == 0 and g[4, 0] == 0 and g[8, 0] == 0
    assert g[1, 0] == -g[3, 0] != 0 and g[2, 0] == -g[6, 0] != 0 and g[5, 0] == -g[7, 0] != 0


def test_host_path_pageable_pinned_and_unstaged_give_the_same_bits(cuda_ctx):
    """The host entry point stages pageable arrays through pinned slots, uploads the caller's initial outputs while the kernel
    runs and adds the FP64 sums afterwards; pinned arrays skip the staging; staging can be switched off. All of that is
    plumbing: the outputs (accumulated onto non-zero initial values) must be the same bits in every mode, and equal to the
    device-resident path with its fused read-modify-write."""
    import torch
    from omega3d_b200.device import DeviceBiotSavart
    n = 300000                 # 1.2 MB per array: several staging decisions per call, ragged tail
    x, s, r = W.random_cloud(n, seed=61)
    nt = 150001
    tx, tr = np.ascontiguousarray(x[:, :nt]), r[:nt].copy()
    rng = np.random.Generator(np.random.MT19937(62))
    u0 = rng.standard_normal((3, nt)).astype(f32)
    g0 = rng.standard_normal((9, nt)).astype(f32)
    outs = []
    for mode in ("pageable", "unstaged", "pinned"):
        cuda_ctx.set_host_staging(mode != "unstaged")
        if mode == "pinned":
            keep = [torch.from_numpy(a.copy()).pin_memory() for a in (x, s, r, tx, tr, u0, g0)]
            ax, as_, ar, atx, atr, u, g = [k.numpy() for k in keep]
        else:
            ax, as_, ar, atx, atr, u, g = x, s, r, tx, tr, u0.copy(), g0.copy()
        cuda_ctx.pts_on_pts(ax, ar, as_, atx, atr, u, g)
        outs.append((u.copy(), g.copy()))
    cuda_ctx.set_host_staging(True)
    for u, g in outs[1:]:
        assert np.array_equal(u.view(np.uint32), outs[0][0].view(np.uint32)) and np.array_equal(g.view(np.uint32), outs[0][1].view(np.uint32))
    # device-resident path: fused read-modify-write in the kernel's epilogue
    eng = DeviceBiotSavart(0, cuda_ctx)
    dev = eng.device
    packed = eng.pack(torch.from_numpy(x).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(r).to(dev))
    du, dg = torch.from_numpy(u0).to(dev), torch.from_numpy(g0).to(dev)
    eng.pts_on_pts(packed, torch.from_numpy(tx).to(dev), torch.from_numpy(tr).to(dev), du, dg)
    torch.cuda.synchronize()
    assert np.array_equal(du.cpu().numpy().view(np.uint32), outs[0][0].view(np.uint32))
    assert np.array_equal(dg.cpu().numpy().view(np.uint32), outs[0][1].view(np.uint32))
    # velocity only, singular targets, through the same plumbing
    u1, u2 = u0.copy(), u0.copy()
    cuda_ctx.pts_on_pts(x, r, s, tx, None, u1, None)
    cuda_ctx.set_host_staging(False)
    cuda_ctx.pts_on_pts(x, r, s, tx, None, u2, None)
    cuda_ctx.set_host_staging(True)
    assert np.array_equal(u1.view(np.uint32), u2.view(np.uint32)) and not np.array_equal(u1, u0)


def test_two_device_context_equals_one_device(cuda_ctx):
    """In-process multi-GPU (one context driving 2 GPUs): targets are partitioned across the devices, sources
    replicated; every target's sum is computed by exactly one device."""
    if cuda_ctx.lib.o3d_cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx2 = I.CudaContext((0, 1))
    n = 60001
    x, s, r = W.random_cloud(n, seed=77)
    u1, g1 = np.zeros((3, n), f32), np.zeros((9, n), f32)
    u2, g2 = np.zeros((3, n), f32), np.zeros((9, n), f32)
    cuda_ctx.pts_on_pts(x, r, s, x, r, u1, g1)
    ctx2.pts_on_pts(x, r, s, x, r, u2, g2)
    assert rel_err(u2, u1) <= 2e-7 and rel_err(g2, g1) <= 2e-7
    nodes, idx = W.icosphere(2, 0.5)
    surf = I.Surfaces(soa(nodes), idx, W.panel_strengths(idx.shape[0], seed=5), I.active)
    a, b = np.zeros((3, n), f32), np.zeros((3, n), f32)
    cuda_ctx.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], x, a, None)
    ctx2.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], x, b, None)
    assert rel_err(b, a) <= 2e-7
    pa, pb = np.zeros((3, surf.np_), f32), np.zeros((3, surf.np_), f32)
    cuda_ctx.pts_on_pan(x, s, surf.x, surf.idx, surf.area, pa)
    ctx2.pts_on_pan(x, s, surf.x, surf.idx, surf.area, pb)
    assert rel_err(pb, pa) <= 2e-7
    assert np.array_equal(I.panels_on_panels_coeff(surf, surf, cuda_ctx), I.panels_on_panels_coeff(surf, surf, ctx2))
    ctx2.close()


# ---- the SASS-post-processed copies of pp2_kernel (tools/sass_patch.py) against the copies as compiled -------------
@pytest.mark.parametrize("ns,nt", [(3, 5), (1000, 777), (4099, 2050), (70000, 40000)])
@pytest.mark.parametrize("grad", [True, False])
@pytest.mark.parametrize("uniform", [True, False])
def test_tuned_kernels_bit_identical(cuda_ctx, ns, nt, grad, uniform):
    """The post-pass only sets operand-reuse bits and clears yield hints: every output bit must be the same."""
    sx, ss, r0 = W.random_cloud(ns, seed=900 + ns)
    tx, _, _ = W.random_cloud(nt, seed=950 + nt)
    sr = r0 if uniform else W.varied_radii(ns, 901, 0.5 * ns ** (-1 / 3), 2.0 * ns ** (-1 / 3))
    tr = np.full(nt, r0[0], f32) if uniform else W.varied_radii(nt, 902, 0.5 * ns ** (-1 / 3), 2.0 * ns ** (-1 / 3))
    out = []
    assert cuda_ctx.tuned_kernels()
    try:
        for on in (True, False):
            cuda_ctx.set_tuned_kernels(on)
            assert cuda_ctx.tuned_kernels() == on
            u = np.full((3, nt), 0.25, f32)
            g = np.full((9, nt), -0.5, f32) if grad else None
            cuda_ctx.pts_on_pts(sx, sr, ss, tx, tr, u, g)
            out.append((u, g))
    finally:
        cuda_ctx.set_tuned_kernels(True)
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    if grad:
        assert np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))
    assert np.all(np.isfinite(out[0][0]))
