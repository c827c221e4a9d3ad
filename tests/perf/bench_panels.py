#!/usr/bin/env python
"""Secondary measurement (not the headline bench): the panel kernels on the flow_over_sphere geometry
(BASELINE configs[3]) through the host C ABI, next to the reference's CPU routines on a bounded sample.
Prints one JSON line per routine. Usage: python tests/perf/bench_panels.py [levels=2] [particles=1000000] [queue|noqueue]
(noqueue: panels -> points with the per-lane round-1 kernel instead of the warp-level work queue)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omega3d_b200 import influence as I  # noqa: E402
from omega3d_b200 import workloads as W  # noqa: E402
from oracle import oracle_py  # noqa: E402

f32 = np.float32


def main():
    levels = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    nodes, idx = W.icosphere(levels, 0.5)
    val = W.panel_strengths(idx.shape[0], seed=3)
    surf = I.Surfaces(np.ascontiguousarray(nodes.T), idx, val, I.active)
    npan = surf.np_
    # particles in a shell around the body (shed vorticity lives near the wall) plus a wake cloud
    rng = np.random.Generator(np.random.MT19937(5))
    x = (rng.random((3, n), dtype=f32) - f32(0.5)) * f32(3.0)
    k = n // 2
    x[:, :k] = x[:, :k] / np.linalg.norm(x[:, :k], axis=0) * (0.5 + 0.15 * rng.random(k) ** 2)
    x = np.ascontiguousarray(x.astype(f32))
    s = ((rng.random((3, n), dtype=f32) - f32(0.5)) / f32(n)).astype(f32)
    ctx = I.CudaContext((0,))
    queue = not (len(sys.argv) > 3 and sys.argv[3] == "noqueue")
    ctx.set_panel_queue(queue)
    try:
        ref = oracle_py.Reference(fast=True)
    except Exception:
        ref = None
    res = oracle_py.Restatement()
    sel = W.strided_subset(n, 2048)

    def emit(name, pairs, ms, flops, cpu_rate, err):
        print(json.dumps({"routine": name, "panel_queue": queue, "panels": npan, "particles": n, "pairs_per_s": pairs / (ms * 1e-3), "kernel_ms": ms,
                          "gflops_reference_count": flops / (ms * 1e-3) * 1e-9, "cpu_pairs_per_s": cpu_rate,
                          "cpu_cores": res.max_threads(), "max_rel_err_vs_oracle_sample": err}), flush=True)

    for grad in (False, True):
        u = np.zeros((3, n), f32); g = np.zeros((9, n), f32) if grad else None
        ctx.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], x, u, g)   # warm-up
        u[:] = 0
        if grad: g[:] = 0
        ctx.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], x, u, g)
        t = ctx.last_timing()
        tx = np.ascontiguousarray(x[:, sel]); ru = np.zeros((3, sel.size), f32); rg = np.zeros((9, sel.size), f32) if grad else None
        t0 = time.perf_counter(); res.pan_on_pts(surf.x, surf.idx, surf.ts, surf.area, surf.ps[2], tx, ru, rg); dt = time.perf_counter() - t0
        err = float(np.max(np.abs(u[:, sel] - ru)) / np.max(np.abs(ru)))
        if grad: err = max(err, float(np.max(np.abs(g[:, sel] - rg)) / np.max(np.abs(rg))))
        emit("panels_affect_points" + ("+grad" if grad else ""), npan * n, t["kernel_ms"], ctx.flops, npan * sel.size / dt, err)

    pu = np.zeros((3, npan), f32)
    ctx.pts_on_pan(x, s, surf.x, surf.idx, surf.area, pu); pu[:] = 0
    ctx.pts_on_pan(x, s, surf.x, surf.idx, surf.area, pu)
    t = ctx.last_timing()
    ns = min(n, 20000)
    rpu = np.zeros((3, npan), f32); gpu_s = np.zeros((3, npan), f32)
    xs, ss = np.ascontiguousarray(x[:, :ns]), np.ascontiguousarray(s[:, :ns])
    t0 = time.perf_counter(); res.pts_on_pan(xs, ss, surf.x, surf.idx, surf.area, rpu); dt = time.perf_counter() - t0
    ctx.pts_on_pan(xs, ss, surf.x, surf.idx, surf.area, gpu_s)
    emit("points_affect_panels", npan * n, t["kernel_ms"], 0.0, npan * ns / dt, float(np.max(np.abs(gpu_s - rpu)) / np.max(np.abs(rpu))))

    rs = I.Surfaces(np.ascontiguousarray(nodes.T), idx, None, I.reactive)
    a = I.panels_on_panels_coeff(rs, rs, ctx)
    a = I.panels_on_panels_coeff(rs, rs, ctx)
    t = ctx.last_timing()
    cpu_rate, err = None, None
    if npan <= 1300:
        t0 = time.perf_counter()
        b = res.pan_on_pan_coeff(rs.x, rs.idx, rs.b1, rs.b2, rs.area, rs.x, rs.idx, rs.b1, rs.b2, rs.nrm, rs.area, True)
        dt = time.perf_counter() - t0
        cpu_rate, err = npan * npan / dt, float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    emit("panels_on_panels_coeff", npan * npan, t["kernel_ms"], ctx.flops, cpu_rate, err)

    # matrix-free A x with the same blocks
    from omega3d_b200 import bem as B
    op = B.PanelOperator(rs, rs, ctx)
    xv = rng.random(3 * npan, dtype=f32) - f32(0.5)
    op.matvec(xv)
    y = op.matvec(xv)
    t = ctx.last_timing()
    err = float(np.max(np.abs(y - a.reshape(3 * npan, 3 * npan).T.astype(np.float64) @ xv)) / np.max(np.abs(y)))
    emit("bem_operator_matvec", npan * npan, t["kernel_ms"], op.flops, None, err)

    # particle x panel closest-point loops: 149 flops per pair by the reference's count
    from omega3d_b200 import reflect as R
    for name, fn, ofn in (("reflect_panp2", lambda p_: R.reflect_panp2(rs, p_, ctx), lambda xx: res.reflect(rs.x, rs.idx, rs.nrm, xx)),
                          ("clear_inner_panp2", lambda p_: R.clear_inner_panp2(1, rs, p_, 0.2, 0.05, ctx),
                           lambda xx: res.clear_inner(rs.x, rs.idx, rs.nrm, xx, 0.2, 0.05))):
        pts = I.Points(x.copy(), s, 0.05, I.active, I.lagrangian)
        fn(I.Points(x.copy(), s, 0.05, I.active, I.lagrangian))   # warm-up
        moved = fn(pts)
        t = ctx.last_timing()
        xs = np.ascontiguousarray(x[:, sel])
        t0 = time.perf_counter(); ofn(xs); dt = time.perf_counter() - t0
        same = bool(np.array_equal(pts.x[:, sel], xs))
        print(json.dumps({"routine": name, "panel_queue": queue, "panels": npan, "particles": n, "pairs_per_s": npan * n / (t["kernel_ms"] * 1e-3),
                          "kernel_ms": t["kernel_ms"], "gflops_reference_count": 149.0 * npan * n / (t["kernel_ms"] * 1e-3) * 1e-9,
                          "cpu_pairs_per_s": npan * sel.size / dt, "cpu_cores": res.max_threads(), "moved": moved,
                          "bit_identical_to_oracle_on_sample": same}), flush=True)


if __name__ == "__main__":
    main()
