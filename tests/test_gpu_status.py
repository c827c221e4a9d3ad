"""GPU (-m gpu): status-file quantities from RESIDENT particles: o3d_cuda_particles_totals against the
reference's own Points<float>::get_total_circ / get_total_impulse (tests/golden/status.npz, minted through oracle/_ref), the
invariants SURVEY.md section 4 proposes as integration checks (circulation and impulse of inviscid rings over many steps), and
the status line written from the device state."""
import numpy as np
import pytest

from conftest import golden
from omega3d_b200 import convection as C
from omega3d_b200 import status as S
from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32


def terms64(x, s):
    """The float32 per-particle terms the reference forms, summed exactly (float64): what the device's FP64 tree must return."""
    x, s = x.astype(f32), s.astype(f32)
    t = np.stack([s[1] * x[2] - s[2] * x[1], s[2] * x[0] - s[0] * x[2], s[0] * x[1] - s[1] * x[0]]).astype(f32)
    return s.astype(np.float64).sum(axis=1), t.astype(np.float64).sum(axis=1), np.abs(s).astype(np.float64).sum(), np.abs(t).astype(np.float64).sum()


@pytest.mark.parametrize("name", ["ring", "leap", "cloud"])
def test_totals_vs_reference(cuda_ctx, name):
    g = golden("status.npz")
    x, s = g[name + "_x"], g[name + "_s"]
    n = x.shape[1]
    d = C.DeviceParticles(cuda_ctx).upload(x, s, np.full(n, 0.05, f32))
    circ, imp = S.totals(d)
    d.close()
    c64, i64, sabs, tabs = terms64(x, s)
    # the device sums the reference's float terms in FP64: exact to double rounding
    assert np.max(np.abs(np.array(circ) - c64)) <= 1e-13 * sabs and np.max(np.abs(np.array(imp) - i64)) <= 1e-13 * tabs
    # the reference: circulation is a sequential DOUBLE sum rounded once to float (1e-6 stated; it is 6e-8 of the result) ...
    assert np.max(np.abs(np.array(circ, f32) - g[name + "_circ"])) <= 1e-6 * max(np.max(np.abs(g[name + "_circ"])), 1e-7 * sabs)
    # ... impulse a sequential FLOAT sum: its own rounding error grows like sqrt(n) eps of the summed magnitudes
    tol = max(1e-6, 4.0 * np.sqrt(n) * 6e-8)
    assert np.max(np.abs(np.array(imp) - g[name + "_imp"].astype(np.float64))) <= tol * tabs


def test_totals_of_a_two_device_collection_equal_one_device(cuda_ctx):
    from omega3d_b200 import influence as I
    if cuda_ctx.lib.o3d_cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = golden("status.npz")
    x, s = g["cloud_x"], g["cloud_s"]
    ctx2 = I.CudaContext((0, 1))
    a = S.totals(C.DeviceParticles(cuda_ctx).upload(x, s, np.full(x.shape[1], 0.05, f32)))
    b = S.totals(C.DeviceParticles(ctx2).upload(x, s, np.full(x.shape[1], 0.05, f32)))
    assert np.allclose(a[0], b[0], rtol=0, atol=1e-12) and np.allclose(a[1], b[1], rtol=0, atol=1e-12)
    ctx2.close()


def test_invariants_of_the_thick_leapfrogging_rings_over_100_steps(cuda_ctx):
    """100 resident RK2 steps (Convection::advect order 2) of the leapfrogging rings with thick cores (16 800 particles):
    total circulation of closed rings is zero and stays zero (to 1e-6 of sum|s|), total impulse is conserved to 0.2 % -
    and both land where the reference's own 100 steps land (golden: its Points methods + its influence templates)."""
    g = golden("status.npz")
    x, s, r = g["leap_x"], g["leap_s"], g["leap_r"]
    dt, steps = float(g["leap_dt"]), int(g["leap_steps"])
    d = C.DeviceParticles(cuda_ctx).upload(x, s, r)
    c0, i0 = S.totals(d)
    d.advect(2, 0.0, dt, (0.0, 0.0, 0.0), steps)
    c1, i1 = S.totals(d)
    d.close()
    sabs = float(np.abs(s).sum())
    inorm = float(np.max(np.abs(i0)))
    assert np.max(np.abs(c0)) <= 1e-7 * sabs and np.max(np.abs(c1)) <= 1e-6 * sabs
    drift = np.max(np.abs(np.array(i1) - np.array(i0))) / inorm
    ref_drift = np.max(np.abs(g["leap_imp_after"].astype(np.float64) - g["leap_imp"])) / inorm
    print(f"\n  impulse drift over {steps} steps: device {drift:.3e}, reference {ref_drift:.3e}; |circ| {np.max(np.abs(c1)):.2e} of sum|s| {sabs:.3f}")
    assert drift <= 2e-3
    assert np.max(np.abs(np.array(i1) - g["leap_imp_after"])) <= 1e-5 * inorm
    assert np.max(np.abs(np.array(c1) - g["leap_circ_after"])) <= 1e-6 * sabs


def test_single_ring_totals_after_10_steps_vs_reference(cuda_ctx):
    """configs[0] as shipped (210 particles on one circle): 10 steps, while the thin ring is still well conditioned."""
    g = golden("status.npz")
    x, s, r = g["ring_x"], g["ring_s"], g["ring_r"]
    d = C.DeviceParticles(cuda_ctx).upload(x, s, r)
    d.advect(2, 0.0, float(g["ring_dt"]), (0.0, 0.0, 0.0), int(g["ring_steps"]))
    c1, i1 = S.totals(d)
    d.close()
    sabs = float(np.abs(s).sum())
    assert np.max(np.abs(np.array(c1) - g["ring_circ_after"])) <= 2e-5 * sabs
    assert np.max(np.abs(np.array(i1) - g["ring_imp_after"])) <= 2e-5 * float(np.max(np.abs(g["ring_imp"])))


def test_status_lines_from_resident_particles(cuda_ctx, tmp_path):
    """Simulation::dump_stats_to_status on the device state: time, Nv, total circulation, and the one-sided time derivative of
    the total impulse (calculate_simple_forces: zero impulse assumed one step before time 0). The expected file is formed here
    from o3d_cuda_particles_totals and printf's %g - the format the reference's operator<< produces (tests/test_status.py pins
    the writer against the reference's own bytes)."""
    x, s, r, dt, fs = W.example_case("leapfrog_vortex_rings_nv", minrad=0.05, ips=0.03)
    d = C.DeviceParticles(cuda_ctx).upload(x, s, r)
    path = str(tmp_path / "run.dat")
    sf = S.StatusFile()
    sf.set_filename(path)
    expect = ["# time Nv gx gy gz fx fy fz"]
    last_t, last_i = -dt, np.zeros(3, f32)
    for k in range(4):
        t = k * dt
        circ, imp = S.totals(d)
        S.dump_stats_to_status(d, sf, t, dt)
        now = np.array(imp, np.float64).astype(f32)
        force = ((now - last_i).astype(np.float64) / (t - last_t)).astype(f32)
        last_t, last_i = t, now
        vals = [f32(t), None] + [f32(c) for c in circ] + list(force)
        expect.append(" ".join((str(d.n) if v is None else "%g" % float(v)) for v in vals))
        d.advect(2, t, dt, fs, 1)
    sf.close()
    d.close()
    got = open(path).read().split("\n")
    assert got[-1] == "" and got[:-1] == expect
    assert abs(float(got[1].split()[5])) > 10.0        # first line: the whole impulse "appears" within one step
