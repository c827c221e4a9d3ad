"""GPU (-m gpu): the drop-in itself. oracle/_ref/libo3d_dropin.so is the REFERENCE's own Influence.h /
Coefficients.h / ExecEnv.h with integration/omega3d_use_cuda.patch applied, compiled with -DUSE_CUDA around the
reference's real Points<float> / Surfaces<float> containers (built in the container that has /root/reference;
the .so travels to the GPU box). Calling the reference's routines with ExecEnv(..., gpu_cuda) must give what the
same routines give with ExecEnv(..., cpu_x86)."""
import os

import numpy as np
import pytest

from conftest import GRAD_TOL, VEL_TOL, golden, rel_err

from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32
CPU_X86, GPU_CUDA = 1, 4


@pytest.fixture(scope="module")
def dropin():
    from oracle import oracle_py
    if not os.path.exists(os.path.join(oracle_py.OUT, "libo3d_dropin.so")):
        pytest.skip("oracle/_ref/libo3d_dropin.so not built (needs /root/reference at build time)")
    lib = oracle_py.Reference(dropin=True)
    assert lib.built_with_cuda()
    return lib


def both(lib, fn):
    out = []
    for accel in (CPU_X86, GPU_CUDA):
        lib.set_accel(accel)
        out.append(fn())
    lib.set_accel(CPU_X86)
    return out


@pytest.mark.parametrize("kind,grad", [("blob", True), ("blob", False), ("field", True), ("tracer", False)])
def test_points_affect_points_through_reference_dispatch(dropin, kind, grad):
    n = 3000
    x, s, r = W.random_cloud(n, seed=61)
    tx, _, tr = W.random_cloud(1111, seed=62)
    tk = {"blob": dropin.TARG_BLOB, "field": dropin.TARG_FIELD, "tracer": dropin.TARG_TRACER}[kind]

    def run():
        tu = np.full((3, 1111), 0.25, f32)            # non-zero start: the arm must accumulate
        tug = np.full((9, 1111), -0.5, f32) if grad else None
        dropin.pts_on_pts(x, r, s, tx, tr if kind == "blob" else None, tu, tug, targ_kind=tk)
        return tu, tug
    (cu, cg), (gu, gg) = both(dropin, run)
    assert rel_err(gu - 0.25, cu - 0.25) <= VEL_TOL
    if grad:
        assert rel_err(gg + 0.5, cg + 0.5) <= GRAD_TOL


def test_panel_routines_through_reference_dispatch(dropin):
    g = golden("panels_80.npz")

    def pan_pts():
        tu, tug = g["u0"].copy(), g["g0"].copy()
        dropin.pan_on_pts(g["nodes_i"], g["idx"], g["val"], g["tx"], None, tu, tug, targ_kind=dropin.TARG_FIELD)
        return tu, tug
    (cu, cg), (gu, gg) = both(dropin, pan_pts)
    assert np.array_equal(cu, g["u_grad"])           # the CPU arm of the patched build is still the reference
    assert rel_err(gu - g["u0"], cu - g["u0"]) <= VEL_TOL and rel_err(gg - g["g0"], cg - g["g0"]) <= GRAD_TOL

    def pts_pan():
        pu = g["pu0"].copy()
        dropin.pts_on_pan(g["psx"], g["psr"], g["pss"], g["nodes_i"], g["idx"], g["val"], pu)
        return pu
    cpu, gpu = both(dropin, pts_pan)
    assert np.array_equal(cpu, g["pu"]) and rel_err(gpu - g["pu0"], cpu - g["pu0"]) <= VEL_TOL

    def pan_pan():
        return dropin.pan_on_pan(g["nodes_i"], g["idx"], g["val"], g["nodes_i"], g["idx"], np.zeros_like(g["val"]))
    cpu, gpu = both(dropin, pan_pan)
    assert np.array_equal(cpu, g["pan_on_pan_pu"]) and rel_err(gpu, cpu) <= VEL_TOL


def test_coefficients_through_reference_dispatch(dropin):
    """-DUSE_CUDA makes the patched panels_on_panels_coeff build its block on the GPU; compare with the golden block."""
    g = golden("coeff_20.npz")
    bc = np.zeros((20, 3), f32)
    a = dropin.pan_on_pan_coeff(g["n0"], g["i0"], bc)
    assert rel_err(a, g["a_self"]) <= 2e-5
    a = dropin.pan_on_pan_coeff(g["n0"], g["i0"], bc, target=(g["n1"], g["i1"], bc))
    assert rel_err(a, g["a_cross"]) <= 2e-5


@pytest.mark.parametrize("order", [1, 2, 3])
def test_convection_through_reference_points(dropin, order):
    """The reference's real Points<float> advanced three ways: (a) its own CPU code, (b) its own host Runge-Kutta
    sequencing with only the influence sums on the GPU (the Influence.h arm), (c) the whole step on the device
    (the Convection.h arm, integration/O3DCudaConvection.h). (b) and (c) must agree bit for bit - the same kernels see
    the same inputs and the device's O(N) kernels round as the reference's do - and both must match (a) within tolerance."""
    g = golden("convection.npz")
    n = g["adv_x"].shape[1]

    def run():
        x, s, e = g["adv_x"].copy(), g["adv_s"].copy(), np.ones(n, f32)
        u, ug = dropin.advect(order, 2, float(g["adv_dt"]), g["adv_fs"], x, s, g["adv_r"], e)
        return x, s, e, u, ug
    dropin.lib.o3d_ref_set_device_convect(0)
    cpu, gpu_sums = both(dropin, run)
    dropin.lib.o3d_ref_set_device_convect(1)
    dropin.set_accel(GPU_CUDA)
    gpu_step = run()
    dropin.set_accel(CPU_X86)
    for a, b in zip(cpu, [g[f"adv{order}_{k}"] for k in ("x", "s", "elong", "u", "ug")]):
        assert np.array_equal(a, b)                  # the CPU arm of the patched build is still the reference
    for a, b, name in zip(gpu_sums, gpu_step, ("x", "s", "elong", "u", "ug")):
        assert np.array_equal(a, b), name
    tol = {"x": 1e-6, "s": 2e-5, "elong": 1e-5, "u": VEL_TOL, "ug": GRAD_TOL}
    for a, b, name in zip(gpu_step, cpu, ("x", "s", "elong", "u", "ug")):
        assert rel_err(a, b) <= tol[name], name


@pytest.mark.parametrize("order", [1, 2])
def test_body_step_through_the_patched_dispatch(dropin, order):
    """Convection::advect around one static body (configs[3]'s sphere) as the patched Convection.h dispatches it: with gpu_cuda
    the particles stay resident (o3d::cuda_advect_particles_body) and the reference's solve_bem runs unchanged on the sums
    the device delivers through its own points_affect_panels dispatch; with cpu_x86 the same driver runs the reference's host
    sequencing (find_derivs / move / clear_inner_layer). Only BEM::solve (Eigen) is replaced, by one dense numpy solve on both
    sides, with the reference's own coefficient matrix."""
    import math
    nodes, idx = W.icosphere(2, 0.5)
    np_ = idx.shape[0]
    rng = np.random.Generator(np.random.MT19937(17))
    n = 6000
    d = rng.standard_normal((3, n // 2)); d /= np.linalg.norm(d, axis=0)
    x = np.concatenate([d * (0.5 + 0.12 * rng.random(n // 2) - 0.01),
                        np.stack([0.4 + 1.5 * rng.random(n - n // 2), 0.7 * (rng.random(n - n // 2) - 0.5), 0.7 * (rng.random(n - n // 2) - 0.5)])], axis=1).astype(f32)
    x = np.ascontiguousarray(x)
    s = ((rng.random((3, n)) - 0.5) * (4.0 / n)).astype(f32)
    ips = 0.0894
    r = np.full(n, 1.5 * ips, f32)
    A = np.asarray(dropin.pan_on_pan_coeff(nodes, idx, np.zeros((np_, 3), f32)), np.float64).reshape(3 * np_, 3 * np_).T
    lu = np.linalg.inv(A)

    def run():
        px, ps, pe = x.copy(), s.copy(), np.ones(n, f32)
        solves, ts = dropin.advect_body(order, 0.02, (1.0, 0.0, 0.0), ips, px, ps, r, pe, nodes, idx, lambda rhs: (lu @ rhs.astype(np.float64)).astype(f32))
        return px, ps, pe, solves, ts
    (cx, cs, ce, ck, cts), (gx, gs, ge, gk, gts) = both(dropin, run)
    assert ck == order and gk == order
    assert np.max(np.abs(cts)) > 0 and rel_err(gts, cts) <= 20 * VEL_TOL       # solved strengths (amplified by cond(A))
    assert not np.array_equal(cx, x)
    assert rel_err(gx, cx) <= 1e-6 and rel_err(gs, cs) <= 2e-5 and rel_err(ge, ce) <= 2e-5
    assert np.min(np.linalg.norm(gx, axis=0)) > 0.49                           # the inner layer was cleared on the device


def test_reflect_and_clear_inner_through_reference_functions(dropin):
    """The patched reflect_panp2 / clear_inner_panp2 (src/Reflect.h + integration hunk) take no ExecEnv: in a -DUSE_CUDA
    build the default back end decides, so calling the reference's own functions here runs the CUDA arm. The fixtures
    are what the unpatched reference returned on the CPU - the comparison is bit for bit."""
    if not hasattr(dropin.lib, "o3d_ref_reflect"):
        pytest.skip("stale drop-in build")
    g = golden("reflect.npz")
    x = g["x0"].copy()
    assert dropin.reflect(g["nodes_i"], g["idx"], x) == int(g["reflect_moved"])
    assert np.array_equal(x, g["reflect_x"])
    x = g["x0"].copy()
    n = dropin.clear_inner(1, g["nodes_i"], g["idx"], x, np.full(x.shape[1], 0.03, f32), float(g["clear_cm"]), float(g["clear_ips"]))
    assert n == int(g["clear_moved"]) and np.array_equal(x, g["clear_x"])


def test_cuda_arm_follows_the_reference_core_define():
    """oracle/_ref/libo3d_dropin_exp.so: the patched reference with "#define USE_EXPONENTIAL_KERNEL" active in its
    src/CoreFunc.h. Its gpu_cuda arm must give what ITS cpu_x86 arm gives - and not what the shipped (WL) core gives.
    Run in a process of its own: both drop-in builds define the same inline context holder."""
    import subprocess
    import sys
    from oracle import oracle_py
    if not os.path.exists(os.path.join(oracle_py.OUT, "libo3d_dropin_exp.so")):
        pytest.skip("oracle/_ref/libo3d_dropin_exp.so not built (needs /root/reference at build time)")
    code = """
import numpy as np, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import rel_err
from oracle import oracle_py
from omega3d_b200 import workloads as W
lib = oracle_py.Reference(dropin="exp")
assert lib.built_with_cuda()
x, s, r = W.random_cloud(3000, seed=61)
out = []
for accel in (1, 4):
    lib.set_accel(accel)
    tu, tug = np.full((3, 3000), 0.25, np.float32), np.full((9, 3000), -0.5, np.float32)
    lib.pts_on_pts(x, r, s, x, r, tu, tug)
    out.append((tu - 0.25, tug + 0.5))
wl = oracle_py.Restatement()
wu, wg = np.zeros((3, 3000), np.float32), np.zeros((9, 3000), np.float32)
wl.pts_on_pts(x, r, s, x, r, wu, wg)
print("ERR", rel_err(out[1][0], out[0][0]), rel_err(out[1][1], out[0][1]), rel_err(out[1][1], wg))
""" % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    eu, eg, e_wl = [float(v) for v in r.stdout.split("ERR")[1].split()]
    assert eu <= VEL_TOL and eg <= GRAD_TOL
    assert e_wl > 100 * GRAD_TOL      # the two cores really differ on this cloud
