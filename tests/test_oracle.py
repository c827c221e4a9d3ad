"""CPU: the oracle restatement against the golden vectors minted from the reference's own code
(tests/golden/make_golden.py), bit for bit; and, where the reference build is present, the reference
library against the same vectors (proves the fixtures are reproducible)."""
import numpy as np
import pytest

from conftest import golden

f32 = np.float32


def soa_nodes(nodes_i):
    return np.ascontiguousarray(nodes_i.T)


def test_kat_single_interaction(restate):
    k = golden("kat.npz")
    out = restate.kernel("0v_0bg", k["s7"], k["t4"])
    assert np.array_equal(out, k["kernel_0v_0bg"])
    # the values recorded in SURVEY.md 8c (i) from the survey's own probe of the reference
    np.testing.assert_allclose(out[:3], [-1.50746942, 0.301493943, -5.4268899], rtol=1e-8)
    np.testing.assert_allclose(out[3:], [7.12934256, -4.44080782, 19.6357574, -4.11440372, 1.42586899, -13.6058807,
                                         3.65342999, -11.584466, -8.55521202], rtol=1e-8)


@pytest.mark.parametrize("name", ["near", "far", "mid", "touch"])
@pytest.mark.parametrize("g", ["v", "g"])
def test_kat_recursive_panel_kernel(restate, name, g):
    k = golden("kat.npz")
    out, flops = restate.rkernel(g == "g", k["tri9"], k["str4_" + g], k[f"rk_{name}_{g}_t"], 0.5)
    assert np.array_equal(out, k[f"rk_{name}_{g}_out"])
    assert flops == int(k[f"rk_{name}_{g}_flops"])  # the reference's own flop bookkeeping, leaf for leaf


def test_kat_flops_value_from_survey(restate):
    k = golden("kat.npz")
    assert int(k["rk_near_v_flops"]) == 4555


@pytest.mark.parametrize("variant,blob,grad", [("0bg", True, True), ("0b", True, False), ("0pg", False, True), ("0p", False, False)])
def test_pts_on_pts_bit_identical(restate, variant, blob, grad):
    g = golden("pts_on_pts.npz")
    tu = g["u0"].copy()
    tug = g["g0"].copy() if grad else None
    restate.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], g["tr"] if blob else None, tu, tug)
    assert np.array_equal(tu, g["u_" + variant])
    if grad:
        assert np.array_equal(tug, g["g_" + variant])


def test_self_cloud_properties(restate):
    g = golden("self_cloud_1000.npz")
    tu, tug = np.zeros((3, 1000), f32), np.zeros((9, 1000), f32)
    restate.pts_on_pts(g["x"], g["r"], g["s"], g["x"], g["r"], tu, tug)
    assert np.array_equal(tu, g["u"]) and np.array_equal(tug, g["g"])
    # vortex-only sources: the velocity gradient is trace-free up to rounding (d . (w x d) = 0)
    trace = tug[0] + tug[4] + tug[8]
    assert np.max(np.abs(trace)) < 1e-4 * np.max(np.abs(tug))


def test_pan_on_pts_bit_identical(restate):
    g = golden("panels_80.npz")
    nodes = soa_nodes(g["nodes_i"])
    sss = np.ascontiguousarray(g["val"][:, 2])
    tu, tug = g["u0"].copy(), g["g0"].copy()
    restate.pan_on_pts(nodes, g["idx"], g["ts"], g["area"], sss, g["tx"], tu, tug)
    assert np.array_equal(tu, g["u_grad"]) and np.array_equal(tug, g["g_grad"])
    tu = g["u0"].copy()
    restate.pan_on_pts(nodes, g["idx"], g["ts"], g["area"], sss, g["tx"], tu, None)
    assert np.array_equal(tu, g["u_vel"])


def test_pts_on_pan_bit_identical(restate):
    g = golden("panels_80.npz")
    pu = g["pu0"].copy()
    restate.pts_on_pan(g["psx"], g["pss"], soa_nodes(g["nodes_i"]), g["idx"], g["area"], pu)
    assert np.array_equal(pu, g["pu"])


def test_coeff_bit_identical(restate):
    g = golden("coeff_20.npz")
    n0, n1 = soa_nodes(g["n0"]), soa_nodes(g["n1"])
    a = restate.pan_on_pan_coeff(n0, g["i0"], g["sb1"], g["sb2"], g["area0"], n0, g["i0"], g["sb1"], g["sb2"], g["snrm"],
                                 g["area0"], True)
    assert np.array_equal(a, g["a_self"])
    a = restate.pan_on_pan_coeff(n0, g["i0"], g["sb1"], g["sb2"], g["area0"], n1, g["i1"], g["tb1"], g["tb2"], g["tnrm"],
                                 g["area1"], False)
    assert np.array_equal(a, g["a_cross"])


def test_reference_reproduces_golden(reference_lib):
    """Only where oracle/_ref/libo3d_ref.so exists: the fixtures are what the reference's code returns."""
    g = golden("pts_on_pts.npz")
    tu, tug = g["u0"].copy(), g["g0"].copy()
    reference_lib.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], g["tr"], tu, tug)
    assert np.array_equal(tu, g["u_0bg"]) and np.array_equal(tug, g["g_0bg"])
    p = golden("panels_80.npz")
    tu = p["u0"].copy()
    reference_lib.pan_on_pts(p["nodes_i"], p["idx"], p["val"], p["tx"], None, tu, None, targ_kind=reference_lib.TARG_TRACER)
    assert np.array_equal(tu, p["u_vel"])
    with pytest.raises(RuntimeError):  # inert lagrangian tracers + velandgrad: the reference asserts (src/Influence.h:368)
        reference_lib.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], None, tu, tug, targ_kind=reference_lib.TARG_TRACER)


# ---- the alternate core functions of src/CoreFunc.h -----------------------------------------------------------------
@pytest.mark.parametrize("core,tag", [(1, "rm"), (2, "exp"), (3, "v2")])
def test_restatement_matches_reference_core_builds_bit_for_bit(restate, core, tag):
    """tests/golden/cores.npz holds outputs of the reference built with USE_RM_KERNEL / USE_EXPONENTIAL_KERNEL /
    USE_V2_KERNEL active in src/CoreFunc.h; the restatement's core argument reproduces each bit for bit."""
    g = golden("cores.npz")
    for variant, blob, grad in (("0bg", True, True), ("0b", True, False), ("0pg", False, True), ("0p", False, False)):
        tu = g["u0"].copy()
        tug = g["g0"].copy() if grad else None
        restate.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], g["tr"] if blob else None, tu, tug, core=core)
        assert np.array_equal(tu, g[f"{tag}_u_{variant}"])
        if grad:
            assert np.array_equal(tug, g[f"{tag}_g_{variant}"])
    u, ug = np.zeros((3, 1000), np.float32), np.zeros((9, 1000), np.float32)
    restate.pts_on_pts(g["cx"], g["cr"], g["cs"], g["cx"], g["cr"], u, ug, core=core)
    assert np.array_equal(u, g[f"{tag}_cloud_u"]) and np.array_equal(ug, g[f"{tag}_cloud_g"])


def test_exponential_core_golden_walks_all_three_branches():
    """The close-pair ladder of cores.npz puts pairs in each arm of exp_cond (src/CoreFunc.h:114-128)."""
    g = golden("cores.npz")
    d = (g["tx"][:, 40:100] - g["sx"][:, 40:100]).astype(np.float64)
    dist = np.sqrt((d * d).sum(0))
    reld3 = dist ** 3 / (g["sr"][40:100].astype(np.float64) ** 3 + g["tr"][40:100].astype(np.float64) ** 3)
    assert (reld3 < 0.001).any() and (reld3 > 16).any() and ((reld3 > 0.001) & (reld3 < 16)).any()


def test_core_zero_is_the_shipped_restatement(restate):
    from oracle.oracle_py import _p
    g = golden("pts_on_pts.npz")
    a, ag = g["u0"].copy(), g["g0"].copy()
    sx, ss, tx = [np.ascontiguousarray(g[k]) for k in ("sx", "ss", "tx")]
    sr, tr = np.ascontiguousarray(g["sr"]), np.ascontiguousarray(g["tr"])
    restate.lib.o3d_oracle_pts_on_pts_core(0, sx.shape[1], _p(sx[0]), _p(sx[1]), _p(sx[2]), _p(sr), _p(ss[0]), _p(ss[1]), _p(ss[2]),
                                           tx.shape[1], _p(tx[0]), _p(tx[1]), _p(tx[2]), _p(tr), _p(a), _p(ag))
    assert np.array_equal(a, g["u_0bg"]) and np.array_equal(ag, g["g_0bg"])


@pytest.mark.parametrize("core,tag", [(1, "rm"), (2, "exp"), (3, "v2")])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_restatement_advect_matches_reference_core_builds(restate, core, tag, order):
    g = golden("cores.npz")
    x, s, e = g["adv_x0"].copy(), g["adv_s0"].copy(), np.ones(300, np.float32)
    restate.advect(order, 2, 0.02, (0.1, 0.0, 0.0), x, s, g["adv_r"], e, core=core)
    assert np.array_equal(x, g[f"{tag}_adv{order}_x"]) and np.array_equal(s, g[f"{tag}_adv{order}_s"])
    assert np.array_equal(e, g[f"{tag}_adv{order}_elong"])


@pytest.mark.parametrize("core,tag", [(1, "rm"), (2, "exp"), (3, "v2")])
def test_reference_core_builds_reproduce_cores_golden(core, tag):
    """Where the core builds of the reference are present (oracle/_ref/libo3d_ref_{rm,exp,v2}.so: built in the container
    that has /root/reference, shipped prebuilt to the GPU box), they reproduce tests/golden/cores.npz bit for bit - the
    fixtures are reproducible outputs of the reference's own code, not of the restatement."""
    import os
    from oracle import oracle_py
    if not (os.path.exists(os.path.join(oracle_py.OUT, oracle_py.Reference.CORE_BUILDS[core])) or oracle_py.have_reference()):
        pytest.skip("core builds of the reference not present")
    ref = oracle_py.Reference(core=core)
    g = golden("cores.npz")
    tu, tug = g["u0"].copy(), g["g0"].copy()
    ref.pts_on_pts(g["sx"], g["sr"], g["ss"], g["tx"], g["tr"], tu, tug)
    assert np.array_equal(tu, g[f"{tag}_u_0bg"]) and np.array_equal(tug, g[f"{tag}_g_0bg"])
    x, s, e = g["adv_x0"].copy(), g["adv_s0"].copy(), np.ones(300, np.float32)
    ref.advect(2, 2, 0.02, (0.1, 0.0, 0.0), x, s, g["adv_r"], e)
    assert np.array_equal(x, g[f"{tag}_adv2_x"]) and np.array_equal(s, g[f"{tag}_adv2_s"])


@pytest.mark.parametrize("core", [0, 1, 2, 3])
def test_restatement_properties_hold_for_every_core(restate, core):
    """Size-independent properties the GPU tests lean on, checked on the checker itself: the vortex-only velocity gradient
    is trace free (d . (d x w) = 0), the sums are exactly linear under power-of-two scaling of the strengths, a particle
    induces no velocity on itself, and a rigid translation by a power of two leaves velocities unchanged to rounding."""
    from omega3d_b200 import workloads as W
    n = 400
    x, s, r = W.random_cloud(n, seed=31, radius=0.07)
    s = (s * np.float32(n)).astype(np.float32)
    u, g = np.zeros((3, n), np.float32), np.zeros((9, n), np.float32)
    restate.pts_on_pts(x, r, s, x, r, u, g, core=core)
    assert np.isfinite(u).all() and np.isfinite(g).all()
    trace = g[0].astype(np.float64) + g[4] + g[8]
    assert np.max(np.abs(trace)) <= 2e-6 * np.max(np.abs(g))
    u2, g2 = np.zeros((3, n), np.float32), np.zeros((9, n), np.float32)
    restate.pts_on_pts(x, r, (s * np.float32(0.25)).astype(np.float32), x, r, u2, g2, core=core)
    assert np.array_equal(u2, u * np.float32(0.25)) and np.array_equal(g2, g * np.float32(0.25))
    # one particle on itself: zero velocity; the gradient keeps only the antisymmetric +-w r3 terms (src/Kernels.h:185-191)
    x1, s1, r1 = x[:, :1].copy(), s[:, :1].copy(), r[:1].copy()
    a, b = np.zeros((3, 1), np.float32), np.zeros((9, 1), np.float32)
    restate.pts_on_pts(x1, r1, s1, x1, r1, a, b, core=core)
    assert not a.any() and b[0, 0] == 0 and b[4, 0] == 0 and b[8, 0] == 0
    assert b[1, 0] == -b[3, 0] and b[2, 0] == -b[6, 0] and b[5, 0] == -b[7, 0] and b[1, 0] != 0
    # translation
    xt = (x + np.float32(2.0)).astype(np.float32)
    ut = np.zeros((3, n), np.float32)
    restate.pts_on_pts(xt, r, s, xt, r, ut, None, core=core)
    assert np.max(np.abs(ut - u)) <= 2e-5 * np.max(np.abs(u))
