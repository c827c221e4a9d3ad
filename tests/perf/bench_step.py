"""Convection steps on resident particles: steps/s and interactions/s of Convection::advect (RK2, the reference's
default order) for the example cases and their grown versions, CUDA-graph replay vs launch-by-launch, next to the
reference's own CPU step (oracle/_ref, the reference's Points methods + influence templates) where it is small enough.

  python tests/perf/bench_step.py [--sizes small|all] [--steps K]      -> one JSON line per case
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omega3d_b200 import convection as C  # noqa: E402
from omega3d_b200 import influence as I  # noqa: E402
from omega3d_b200 import workloads as W  # noqa: E402


def cases(which):
    out = [("single_vortex_ring_nv (C1 as shipped)", W.example_case("single_vortex_ring_nv")),
           ("leapfrog_vortex_rings_nv (C2 as shipped)", W.example_case("leapfrog_vortex_rings_nv")),
           ("leapfrog, thick rings minrad 0.05 ips 0.015", W.example_case("leapfrog_vortex_rings_nv", minrad=0.05, ips=0.015))]
    if which == "all":
        out.append(("leapfrog, thick rings minrad 0.1 ips 0.0095 (~260K)", W.example_case("leapfrog_vortex_rings_nv", minrad=0.1, ips=0.0095)))
        out.append(("leapfrog, thick rings minrad 0.06 ips 0.004 (~1M, C2 grown)", W.example_case("leapfrog_vortex_rings_nv", minrad=0.06, ips=0.004)))
    return out


def body_step(ctx, a):
    """Resident RK step with a static body (3Dexamples/flow_over_sphere.json's path): what crosses the host boundary per derivative
    evaluation is 3 np floats out and 4 np floats back (np = 320 panels); the particle arrays stay in HBM."""
    f32 = np.float32
    n, ips, dt, fs = a.body, 0.0894, 0.02, (1.0, 0.0, 0.0)
    nodes, idx = W.icosphere(2, 0.5)
    surf = I.Surfaces(np.ascontiguousarray(nodes.T), idx, None, I.reactive)
    rng = np.random.Generator(np.random.MT19937(3))
    m = n // 2
    d = rng.standard_normal((3, m)); d /= np.linalg.norm(d, axis=0)
    shell = d * (0.5 + 0.12 * rng.random(m) - 0.01)
    wake = np.stack([0.4 + 1.6 * rng.random(n - m), 0.7 * (rng.random(n - m) - 0.5), 0.7 * (rng.random(n - m) - 0.5)])
    x = np.ascontiguousarray(np.concatenate([shell, wake], axis=1).astype(f32))
    s = np.ascontiguousarray(((rng.random((3, n)) - 0.5) * (4.0 / n)).astype(f32))
    r = np.full(n, 1.5 * ips, f32)
    calls = [0]
    ts = (0.01 * rng.standard_normal((3, surf.np_))).astype(f32)

    def solve(pu):                      # stands in for the reference's host solve: fixed strengths, counted
        calls[0] += 1
        return ts, None

    p = C.DeviceParticles(ctx).upload(x, s, r)
    p.set_body(surf, ips, solve)
    p.advect(a.order, 0.0, dt, fs, 1)   # warm-up
    steps = a.steps or 2
    calls[0] = 0
    t0 = time.perf_counter()
    p.advect(a.order, 0.0, dt, fs, steps)
    wall = time.perf_counter() - t0
    tm = ctx.last_timing()
    moved, solves = p.body_counters()
    print(json.dumps({"case": "resident RK step with a 320-panel sphere attached (C4 path)", "particles": n, "panels": int(surf.np_), "order": a.order,
                      "steps": steps, "wall_ms_per_step": wall * 1e3 / steps, "kernel_ms_per_step": tm["kernel_ms"] / steps,
                      "h2d_ms_per_step": tm["h2d_ms"] / steps, "d2h_ms_per_step": tm["d2h_ms"] / steps, "launches_per_step": tm["launches"] / steps,
                      "bem_callbacks_per_step": calls[0] / steps, "host_floats_out_per_callback": 3 * int(surf.np_), "host_floats_back_per_callback": 4 * int(surf.np_),
                      "particles_moved_out_of_the_inner_layer": int(moved), "solves": int(solves),
                      "particle_particle_interactions_per_s": a.order * float(n) * n / (wall / steps)}), flush=True)
    p.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="small")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--body", type=int, default=0, help="N > 0: instead of the ring cases, one RK step of N particles around a 320-panel sphere "
                    "attached to the resident collection (config C4): BEM right-hand side on the device + host callback per derivative "
                    "evaluation, panels -> particles in find_vels, clear-inner after every move")
    a = ap.parse_args()
    ctx = I.CudaContext((0,))
    if a.body:
        return body_step(ctx, a)
    ref = None
    if not a.no_cpu:
        from oracle import oracle_py
        try:
            ref = oracle_py.Reference(fast=True)
            assert hasattr(ref.lib, "o3d_ref_advect")
        except Exception:
            ref = oracle_py.Restatement()
    for name, (x, s, r, dt, fs) in cases(a.sizes):
        n = x.shape[1]
        steps = a.steps or (200 if n < 20000 else 20 if n < 100000 else 3 if n < 500000 else 1)
        row = {"case": name, "particles": n, "order": a.order, "dt": dt, "steps": steps}
        for graphs in (1, 0):
            ctx.check(ctx.lib.o3d_cuda_set_graphs(ctx.h, graphs))
            p = C.DeviceParticles(ctx).upload(x, s, r)
            p.advect(a.order, 0.0, dt, fs, 2 if n < 500000 else 1)          # warm-up (and the graph capture)
            t0 = time.perf_counter()
            p.advect(a.order, 0.0, dt, fs, steps)
            wall = time.perf_counter() - t0
            tm = ctx.last_timing()
            key = "graph" if graphs else "eager"
            row[key + "_ms_per_step"] = tm["kernel_ms"] / steps
            row[key + "_wall_ms_per_step"] = wall * 1e3 / steps
            row[key + "_launches_per_step"] = tm["launches"] / steps
            if n >= 100000:
                row["interactions_per_s"] = a.order * n * n / (tm["kernel_ms"] * 1e-3 / steps)
                p.close()
                break
            p.close()
        ctx.check(ctx.lib.o3d_cuda_set_graphs(ctx.h, 1))
        if ref is not None and n <= 20000:
            csteps = 20 if n < 1000 else 1
            xx, ss, ee = x.copy(), s.copy(), np.ones(n, np.float32)
            t0 = time.perf_counter()
            ref.advect(a.order, csteps, dt, fs, xx, ss, r, ee)
            row["cpu_reference_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / csteps
            row["cpu_cores"] = ref.max_threads()
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
