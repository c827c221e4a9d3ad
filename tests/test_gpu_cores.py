"""GPU (-m gpu): particles -> points with the alternate core functions of the reference's src/CoreFunc.h
(Rosenhead-Moore, exponential, Vatistas n=2; o3d_cuda_set_core_func), through the C ABI, against

* tests/golden/cores.npz - outputs of three builds of the reference's own templates, each with another
  "#define USE_*_KERNEL" active in src/CoreFunc.h (oracle/Makefile core_build, tests/golden/make_golden.py cores), and
* the plain-C restatement (oracle/biot_oracle.c: o3d_oracle_pts_on_pts_core), bit-identical to those builds
  (tests/test_oracle.py).

Tolerances are the north star's: velocity 1e-5, gradient 1e-4 (max-norm relative).
"""
import numpy as np
import pytest

from conftest import GRAD_TOL, VEL_TOL, golden, rel_err

from omega3d_b200 import influence as I
from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32

CORES = [(I.core_t.rm, "rm"), (I.core_t.exp, "exp"), (I.core_t.v2, "v2")]
VARIANTS = [("0bg", True, True), ("0b", True, False), ("0pg", False, True), ("0p", False, False)]
# flops_t{v,p}_{grads,nograds} of src/CoreFunc.h per core: (tv_grads, tp_grads, tv_nograds, tp_nograds)
CORE_FLOPS = {"wl": (16, 14, 10, 8), "rm": (9, 7, 7, 5), "exp": (14, 11, 12, 9), "v2": (13, 10, 11, 8)}


@pytest.fixture(scope="module")
def ctx():
    """Own context: the core function is context state and the session-wide context must stay on the default."""
    c = I.CudaContext((0,))
    yield c
    c.close()


def pair_flops(tag, blob, grad):
    tvg, tpg, tvn, tpn = CORE_FLOPS[tag]
    return (54 + (tvg if blob else tpg)) if grad else (23 + (tvn if blob else tpn))


def test_default_core_is_winckelmans_leonard(ctx):
    assert ctx.core_func() == I.core_t.wl
    with pytest.raises(I.O3DError):
        ctx.set_core_func(4)
    with pytest.raises(I.O3DError):
        ctx.set_core_func(-1)
    assert ctx.core_func() == I.core_t.wl


@pytest.mark.parametrize("core,tag", CORES)
@pytest.mark.parametrize("variant,blob,grad", VARIANTS)
def test_cores_golden(ctx, core, tag, variant, blob, grad):
    g = golden("cores.npz")
    ctx.set_core_func(core)
    try:
        tu = g["u0"].copy()
        tug = g["g0"].copy() if grad else None
        ctx.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], g["tr"] if blob else None, tu, tug)
    finally:
        ctx.set_core_func("wl")
    assert rel_err(tu, g[f"{tag}_u_{variant}"]) <= VEL_TOL
    if grad:
        assert rel_err(tug, g[f"{tag}_g_{variant}"]) <= GRAD_TOL
    ns, nt = g["sx"].shape[1], g["tx"].shape[1]
    assert ctx.flops == nt * ((12 if grad else 3) + pair_flops(tag, blob, grad) * ns)


@pytest.mark.parametrize("core,tag", CORES)
def test_cores_self_cloud_golden(ctx, core, tag):
    """Uniform radii, sources alias targets: every target meets itself (zero velocity, antisymmetric gradient terms)."""
    g = golden("cores.npz")
    p = I.Points(g["cx"], g["cs"], g["cr"], I.active, I.lagrangian)
    p.zero_vels()
    ctx.set_core_func(core)
    try:
        I.points_affect_points(p, p, I.ResultsType(I.velandgrad), I.ExecEnv(), ctx)
    finally:
        ctx.set_core_func("wl")
    assert np.isfinite(p.u).all() and np.isfinite(p.ug).all()
    assert rel_err(p.u, g[f"{tag}_cloud_u"]) <= VEL_TOL and rel_err(p.ug, g[f"{tag}_cloud_g"]) <= GRAD_TOL


@pytest.mark.parametrize("core,tag", CORES)
@pytest.mark.parametrize("ns,nt", [(1, 1), (2, 3), (511, 513), (4099, 2050), (30000, 777)])
def test_cores_vs_oracle_ragged(ctx, restate, core, tag, ns, nt):
    """Ragged sizes (partial tiles, partial CTAs, the gridDim.y source split for small target counts), per-particle radii,
    += on non-zero outputs."""
    sx, ss, _ = W.random_cloud(ns, seed=100 + ns)
    sr = W.varied_radii(ns, 7 + ns, 0.02, 0.09)
    ss = (ss * f32(ns)).astype(f32)
    tx, _, _ = W.random_cloud(nt, seed=200 + nt)
    tr = W.varied_radii(nt, 9 + nt, 0.01, 0.07)
    rng = np.random.default_rng(ns * 31 + nt)
    u0 = (rng.random((3, nt), dtype=f32) - f32(0.5)).astype(f32)
    g0 = (rng.random((9, nt), dtype=f32) - f32(0.5)).astype(f32)
    for blob, grad in ((True, True), (False, False)):
        tu, tug = u0.copy(), (g0.copy() if grad else None)
        ru, rg = u0.copy(), (g0.copy() if grad else None)
        ctx.set_core_func(core)
        try:
            ctx.pts_on_pts(sx, sr, ss, tx, tr if blob else None, tu, tug)
        finally:
            ctx.set_core_func("wl")
        restate.pts_on_pts(sx, sr, ss, tx, tr if blob else None, ru, rg, core=int(core))
        assert rel_err(tu - u0, ru - u0) <= VEL_TOL
        if grad:
            assert rel_err(tug - g0, rg - g0) <= GRAD_TOL


@pytest.mark.parametrize("core,tag", CORES)
def test_cores_properties_at_256k(ctx, restate, core, tag):
    """At a size the oracle cannot sweep: a strided target sample against the oracle, the trace-free gradient, and exact
    power-of-two linearity in the strengths."""
    n = 1 << 18
    x, s, r = W.random_cloud(n)
    u, ug = np.zeros((3, n), f32), np.zeros((9, n), f32)
    ctx.set_core_func(core)
    try:
        ctx.pts_on_pts(x, r, s, x, r, u, ug)
        u2, g2 = np.zeros((3, n), f32), np.zeros((9, n), f32)
        ctx.pts_on_pts(x, r, (s * f32(4.0)).astype(f32), x, r, u2, g2)
    finally:
        ctx.set_core_func("wl")
    assert np.isfinite(u).all() and np.isfinite(ug).all()
    sel = np.arange(0, n, n // 128)
    tx = np.ascontiguousarray(x[:, sel]); tr = np.ascontiguousarray(r[sel])
    ru, rg = np.zeros((3, sel.size), f32), np.zeros((9, sel.size), f32)
    restate.pts_on_pts(x, r, s, tx, tr, ru, rg, core=int(core))
    assert rel_err(u[:, sel], ru) <= VEL_TOL and rel_err(ug[:, sel], rg) <= GRAD_TOL
    trace = ug[0].astype(np.float64) + ug[4] + ug[8]
    assert np.max(np.abs(trace)) <= 1e-5 * np.max(np.abs(ug))
    assert np.array_equal(u2, u * f32(4.0)) and np.array_equal(g2, ug * f32(4.0))


@pytest.mark.parametrize("core,tag", CORES)
def test_cores_resident_find_vels(ctx, restate, core, tag):
    """The device-resident collection (Convection::find_vels: zero -> pack -> kernel -> finalize) follows the context's core."""
    from omega3d_b200 import convection as C
    n = 3000
    x, s, r = W.random_cloud(n, seed=77, radius=0.06)
    fs = (0.1, -0.2, 0.05)
    ctx.set_core_func(core)
    try:
        d = C.DeviceParticles(ctx).upload(x, s, r)
        d.find_vels(fs)
        out = d.download(("u", "ug"))
        d.close()
    finally:
        ctx.set_core_func("wl")
    ru, rg = np.zeros((3, n), f32), np.zeros((9, n), f32)
    restate.pts_on_pts(x, r, s, x, r, ru, rg, core=int(core))
    restate.finalize_vels(ru, rg, fs)
    assert rel_err(out["u"], ru) <= VEL_TOL and rel_err(out["ug"], rg) <= GRAD_TOL


@pytest.mark.parametrize("core,tag", CORES)
@pytest.mark.parametrize("order", [1, 2, 3])
def test_cores_advect_golden(ctx, core, tag, order):
    """Two Runge-Kutta steps on resident particles (second one replayed from the captured CUDA graph) against the same
    steps through the Points methods of the reference build with that core (tests/golden/cores.npz)."""
    from omega3d_b200 import convection as C
    g = golden("cores.npz")
    ctx.set_core_func(core)
    try:
        d = C.DeviceParticles(ctx).upload(g["adv_x0"], g["adv_s0"], g["adv_r"])
        d.advect(order, 0.0, 0.02, (0.1, 0.0, 0.0), 2)
        out = d.download(("x", "s", "elong"))
        d.close()
    finally:
        ctx.set_core_func("wl")
    assert rel_err(out["x"], g[f"{tag}_adv{order}_x"]) <= 1e-6
    assert rel_err(out["s"], g[f"{tag}_adv{order}_s"]) <= 2e-5
    assert rel_err(out["elong"], g[f"{tag}_adv{order}_elong"]) <= 2e-5


def test_core_change_drops_the_captured_graph(ctx, restate):
    """A resident collection that has captured its step under one core must not replay it under another."""
    from omega3d_b200 import convection as C
    g = golden("cores.npz")
    x0, s0, r = g["adv_x0"], g["adv_s0"], g["adv_r"]
    d = C.DeviceParticles(ctx).upload(x0, s0, r)
    try:
        d.advect(2, 0.0, 0.02, (0.1, 0.0, 0.0), 2)          # Winckelmans-Leonard, graph captured
        ctx.set_core_func("v2")
        d.upload(x0, s0, r)
        d.advect(2, 0.0, 0.02, (0.1, 0.0, 0.0), 2)
        out = d.download(("x", "s"))
    finally:
        ctx.set_core_func("wl")
        d.close()
    assert rel_err(out["x"], g["v2_adv2_x"]) <= 1e-6 and rel_err(out["s"], g["v2_adv2_s"]) <= 2e-5


@pytest.mark.parametrize("core,tag", CORES)
def test_panel_kernels_under_other_cores(ctx, core, tag):
    """Panel leaves evaluate the core function at zero radius, where all four reduce to |d|^-3 (bbb = -3 |d|^-5): the
    panel kernels need no per-core variant. Checked against panels -> points of each core build of the reference; the
    flop figure follows the core's leaf count (flops_0vs_0pg = 79 + flops_tp_grads, src/Kernels.h:294)."""
    g, pg = golden("cores.npz"), golden("panels_80.npz")
    surf = I.Surfaces(np.ascontiguousarray(pg["nodes_i"].T), pg["idx"], pg["val"], I.active, I.fixed)
    fld = I.Points(pg["tx"], e=I.inert, m=I.fixed)
    fld.u[:], fld.ug[:] = pg["u0"], pg["g0"]
    I.panels_affect_points(surf, fld, I.ResultsType(I.velonly), I.ExecEnv(), ctx)
    wl_flops = ctx.flops
    fld.u[:], fld.ug[:] = pg["u0"], pg["g0"]
    ctx.set_core_func(core)
    try:
        I.panels_affect_points(surf, fld, I.ResultsType(I.velonly), I.ExecEnv(), ctx)
        flops = ctx.flops
    finally:
        ctx.set_core_func("wl")
    assert rel_err(fld.u - pg["u0"], g[f"{tag}_pan_u"] - pg["u0"]) <= VEL_TOL
    assert rel_err(fld.ug - pg["g0"], g[f"{tag}_pan_g"] - pg["g0"]) <= GRAD_TOL
    # same traversal, cheaper leaves: the difference is (leaves) x (tp_grads difference), a whole multiple of it
    dl = 14 - CORE_FLOPS[tag][1]
    assert flops < wl_flops and (wl_flops - flops) % dl == 0


@pytest.mark.parametrize("core,tag", CORES)
@pytest.mark.parametrize("grad", [True, False])
def test_cores_tuned_kernels_bit_identical(ctx, core, tag, grad):
    """The alternate-core kernels also run from the SASS-post-processed cubin (operand-reuse bits and yield hints only):
    every output bit equals the copy linked into the library."""
    ns, nt = 4099, 2050
    sx, ss, _ = W.random_cloud(ns, seed=900 + ns)
    tx, _, _ = W.random_cloud(nt, seed=950 + nt)
    sr, tr = W.varied_radii(ns, 901, 0.03, 0.12), W.varied_radii(nt, 902, 0.03, 0.12)
    out = []
    ctx.set_core_func(core)
    try:
        for on in (True, False):
            ctx.set_tuned_kernels(on)
            u = np.full((3, nt), 0.25, f32)
            g = np.full((9, nt), -0.5, f32) if grad else None
            ctx.pts_on_pts(sx, sr, ss, tx, tr, u, g)
            out.append((u, g))
    finally:
        ctx.set_tuned_kernels(True)
        ctx.set_core_func("wl")
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    if grad:
        assert np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))
    assert np.all(np.isfinite(out[0][0]))


def test_switching_back_restores_the_default_kernel(ctx, restate):
    g = golden("self_cloud_1000.npz")
    ctx.set_core_func("v2")
    ctx.set_core_func("wl")
    u, ug = np.zeros((3, 1000), f32), np.zeros((9, 1000), f32)
    ctx.pts_on_pts(g["x"], g["r"], g["s"], g["x"], g["r"], u, ug)
    assert rel_err(u, g["u"]) <= VEL_TOL and rel_err(ug, g["g"]) <= GRAD_TOL
    assert ctx.flops == 1000 * (12 + 70 * 1000)
