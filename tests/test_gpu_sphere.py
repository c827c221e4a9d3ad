"""GPU (-m gpu): BASELINE configs[3] (3Dexamples/flow_over_sphere.json) as a PIPELINE - every routine of the path a step with a
body runs, in the reference's order, against the same sequence through the reference's own templates (oracle/_ref):

  solve_bem (src/BEMHelper.h:44-262):  zero panel velocities -> points_affect_panels -> finalize_vels(fs) -> right-hand side
                                       (src/RHS.h) -> A = panels_on_panels_coeff -> solve -> strengths into the surface
  find_vels (src/Convection.h:130-184): zero -> points_affect_points -> panels_affect_points -> finalize_vels(fs)
  move (src/Points.h:288-351), clear_inner_layer (src/Reflect.h:625-655)

The full application needs Eigen (its GMRES) to link; the dense 960 x 960 system is solved by numpy on BOTH sides instead,
which is the one substitution. Geometry: the 320-panel sphere the input file's body refines to,
freestream (1,0,0), a cloud of shed-like particles around and behind it.
"""
import math

import numpy as np
import pytest

from conftest import GRAD_TOL, VEL_TOL, rel_err
from omega3d_b200 import bem as B
from omega3d_b200 import convection as C
from omega3d_b200 import influence as I
from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32
FS = (1.0, 0.0, 0.0)
IPS, DT = 0.0894, 0.02          # ips = sqrt(8) * sqrt(dt / Re) of the input file (C4)
CUT = 0.5 / math.sqrt(2.0 * math.pi)


def wake_cloud(n, seed=3):
    """Particles in a shell just off the sphere (some inside the layer clear_inner_layer empties) and in a wake behind it."""
    rng = np.random.Generator(np.random.MT19937(seed))
    m = n // 2
    d = rng.standard_normal((3, m))
    d /= np.linalg.norm(d, axis=0)
    shell = d * (0.5 + 0.12 * rng.random(m) - 0.01)
    wake = np.stack([0.4 + 1.6 * rng.random(n - m), 0.7 * (rng.random(n - m) - 0.5), 0.7 * (rng.random(n - m) - 0.5)])
    x = np.concatenate([shell, wake], axis=1).astype(f32)
    s = ((rng.random((3, n)) - 0.5) * (4.0 / n)).astype(f32)
    r = np.full(n, 1.5 * IPS, f32)
    return np.ascontiguousarray(x), np.ascontiguousarray(s), r


@pytest.fixture(scope="module")
def sphere():
    nodes, idx = W.icosphere(2, 0.5)
    surf = I.Surfaces(np.ascontiguousarray(nodes.T), idx, None, I.reactive)
    return nodes, idx, surf


def reference_step(ref, nodes, idx, surf, A, x, s, r, order=1):
    """One Convection::advect (order 1 or 2) with the body, every sum through the reference's templates. Returns the state
    after the step and the intermediate quantities of the FIRST derivative evaluation."""
    np_ = idx.shape[0]
    n = x.shape[1]
    keep = {}

    def find_derivs(px, ps):
        pu = np.zeros((3, np_), f32)
        ref.pts_on_pan(px, r, ps, nodes, idx, np.zeros((np_, 3), f32), pu)
        pu = (np.asarray(FS)[:, None] + pu.astype(np.float64) * (0.25 / math.pi)).astype(f32)     # Surfaces::finalize_vels
        surf.pu[:] = pu
        rhs = B.vels_to_rhs_panels(surf)
        sol = np.linalg.solve(A, rhs.astype(np.float64)).astype(f32)
        val = np.ascontiguousarray(sol.reshape(np_, 3))
        u, g = np.zeros((3, n), f32), np.zeros((9, n), f32)
        ref.pts_on_pts(px, r, ps, px, r, u, g)
        ref.pan_on_pts(nodes, idx, val, px, r, u, g, ref.TARG_BLOB)
        ref.finalize_vels(u, g, FS)
        keep.setdefault("pu", pu); keep.setdefault("sol", sol); keep.setdefault("u", u.copy()); keep.setdefault("g", g.copy())
        return u, g

    x, s, e = x.copy(), s.copy(), np.ones(n, f32)
    u0, g0 = find_derivs(x, s)
    moved = 0
    if order == 1:
        ref.move(1, DT, [1.0], [u0], [g0], x, s, e)
    else:
        xi, si = x.copy(), s.copy()
        ref.move(1, (2.0 / 3.0) * DT, [1.0], [u0], [g0], xi, si, None)
        moved += ref.clear_inner(1, nodes, idx, xi, r, CUT, IPS)
        u1, g1 = find_derivs(xi, si)
        ref.move(2, DT, [0.25, 0.75], [u0, u1], [g0, g1], x, s, e, np.zeros((3, n), f32))
    moved += ref.clear_inner(1, nodes, idx, x, r, CUT, IPS)
    return x, s, e, moved, keep


@pytest.mark.parametrize("order", [1, 2])
def test_flow_over_sphere_step_vs_reference_templates(cuda_ctx, reference_lib, sphere, order):
    nodes, idx, surf = sphere
    np_ = idx.shape[0]
    n = 20000
    x, s, r = wake_cloud(n)
    # the influence matrix: GPU-assembled vs the reference's (column-major (3 np)^2), then ONE matrix for both solves so
    # that the comparison below measures the path, not the conditioning of the system (reported)
    A_gpu = I.panels_on_panels_coeff(surf, surf, cuda_ctx)
    A_ref = reference_lib.pan_on_pan_coeff(nodes, idx, np.zeros((np_, 3), f32))
    assert rel_err(A_gpu, A_ref) <= 2e-5
    A = np.asarray(A_ref, np.float64).reshape(3 * np_, 3 * np_).T
    rx, rs, re, rmoved, keep = reference_step(reference_lib, nodes, idx, surf, A, x, s, r, order)

    class RefMatrixBEM(B.DenseBEM):
        pass
    bem = RefMatrixBEM(A_ref, 3 * np_)
    seen = {}

    def solve(pu):
        seen.setdefault("raw", pu.copy())
        surf.pu[:] = pu
        surf.finalize_vels(FS)
        B.solve_bem_for(surf, bem)
        seen.setdefault("pu", surf.pu.copy()); seen.setdefault("sol", bem.getStrengths().copy())
        return surf.ts, surf.ps[2]

    d = C.DeviceParticles(cuda_ctx).upload(x, s, r)
    d.set_body(surf, IPS, solve)
    # the right-hand-side velocities alone, then the step
    pu_only = d.body_vels()
    d.advect(order, 0.0, DT, FS, 1)
    moved, solves = d.body_counters()
    out = d.download()
    d.close()
    assert solves == order
    # BEM right-hand side and solution of the first evaluation
    assert np.array_equal(pu_only, seen["raw"])
    assert rel_err(seen["pu"], keep["pu"]) <= VEL_TOL
    assert rel_err(seen["sol"], keep["sol"]) <= 20 * VEL_TOL          # the solve amplifies the rhs difference (cond(A) below)
    # the state after the step
    ex, es, ee = rel_err(out["x"], rx), rel_err(out["s"], rs), rel_err(out["elong"], re)
    print(f"\n  order {order}: cond(A) {np.linalg.cond(A):.1f}; rhs {rel_err(seen['pu'], keep['pu']):.2e} strengths "
          f"{rel_err(seen['sol'], keep['sol']):.2e}; after the step: x {ex:.2e} s {es:.2e} elong {ee:.2e}; pushed out {moved} (reference {rmoved})")
    assert moved == rmoved and moved > 0
    assert ex <= 1e-6 and es <= 2e-5 and ee <= 2e-5


def test_find_vels_with_body_vs_reference(cuda_ctx, reference_lib, sphere):
    """Convection::find_vels(fs, vort, bdry, vort) with FIXED panel strengths: particles + panels + freestream on every particle,
    velocity and gradient, against the reference's two routines in its order."""
    nodes, idx, surf = sphere
    np_ = idx.shape[0]
    n = 30000
    x, s, r = wake_cloud(n, seed=8)
    val = W.panel_strengths(np_, seed=4)
    act = I.Surfaces(surf.x, idx, val, I.active)
    d = C.DeviceParticles(cuda_ctx).upload(x, s, r)
    d.set_body(surf, IPS, None)
    d.set_body_strengths(act.ts, act.ps[2])
    d.find_vels(FS)
    out = d.download(("u", "ug"))
    d.close()
    u, g = np.zeros((3, n), f32), np.zeros((9, n), f32)
    reference_lib.pts_on_pts(x, r, s, x, r, u, g)
    reference_lib.pan_on_pts(nodes, idx, val, x, r, u, g, reference_lib.TARG_BLOB)
    reference_lib.finalize_vels(u, g, FS)
    assert rel_err(out["u"], u) <= VEL_TOL and rel_err(out["ug"], g) <= GRAD_TOL


def test_resident_clear_inner_bit_identical(cuda_ctx, reference_lib, sphere):
    nodes, idx, surf = sphere
    x, s, r = wake_cloud(50000, seed=5)
    d = C.DeviceParticles(cuda_ctx).upload(x, s, r)
    d.set_body(surf, IPS, None)
    moved = d.clear_inner()
    out = d.download(("x",))
    d.close()
    rx = x.copy()
    rmoved = reference_lib.clear_inner(1, nodes, idx, rx, r, CUT, IPS)
    assert moved == rmoved and moved > 0
    assert np.array_equal(out["x"].view(np.uint32), rx.view(np.uint32))


def test_convection_mirror_with_body(cuda_ctx, sphere):
    """The Python mirror of Convection::advect with bdry + bem: the GPU-assembled dense system, two steps; body flow must keep
    particles out of the sphere and leave the collection finite."""
    nodes, idx, surf = sphere
    x, s, r = wake_cloud(5000, seed=9)
    pts = I.Points(x, s, r, I.active, I.lagrangian)
    bem = B.DenseBEM(I.panels_on_panels_coeff(surf, surf, cuda_ctx), 3 * surf.np_)
    conv = C.Convection(2, ctx=cuda_ctx)
    conv.advect(0.0, DT, FS, IPS, [pts], [surf], [], bem, nsteps=2)
    assert np.all(np.isfinite(pts.x)) and np.all(np.isfinite(pts.s))
    assert np.min(np.linalg.norm(pts.x, axis=0)) >= 0.5 - 2e-3
    assert np.max(np.abs(surf.ts)) > 0
