#!/usr/bin/env python
"""Secondary measurement (not the headline bench): particles -> points under each core function of the reference's
src/CoreFunc.h (o3d_cuda_set_core_func) through the host C ABI; one JSON line per (core, result type) with the
kernel time, interactions/s, FLOP/s by the reference's own per-core flop count and the error of a strided target
sample against the oracle. Usage: python tests/perf/bench_cores.py [particles=262144]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omega3d_b200 import influence as I  # noqa: E402
from omega3d_b200 import workloads as W  # noqa: E402
from oracle import oracle_py  # noqa: E402

f32 = np.float32


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    x, s, r = W.random_cloud(n)
    ctx = I.CudaContext((0,))
    props = ctx.device_props(0)
    res = oracle_py.Restatement()
    sel = W.strided_subset(n, 256)
    tx, tr = np.ascontiguousarray(x[:, sel]), np.ascontiguousarray(r[sel])
    peak = None
    try:
        peak = props["sm_count"] * 128 * 2 * props["clock_khz"] * 1e3
    except Exception:
        pass
    for core in I.core_t:
        ctx.set_core_func(core)
        for grad in (True, False):
            best = 1e30
            for _ in range(3):
                u = np.zeros((3, n), f32)
                g = np.zeros((9, n), f32) if grad else None
                ctx.pts_on_pts(x, r, s, x, r, u, g)
                best = min(best, ctx.last_timing()["kernel_ms"])
            ru = np.zeros((3, sel.size), f32)
            rg = np.zeros((9, sel.size), f32) if grad else None
            res.pts_on_pts(x, r, s, tx, tr, ru, rg, core=int(core))
            eu = float(np.max(np.abs(u[:, sel] - ru)) / np.max(np.abs(ru)))
            eg = float(np.max(np.abs(g[:, sel] - rg)) / np.max(np.abs(rg))) if grad else None
            line = {"core": core.name, "results": "velandgrad" if grad else "velonly", "particles": n, "kernel_ms": best,
                    "interactions_per_s": float(n) * n / (best * 1e-3), "flops_reference_count": ctx.flops,
                    "tflops_reference_count": ctx.flops / (best * 1e-3) * 1e-12, "vel_err": eu, "grad_err": eg}
            if peak:
                line["frac_fp32_peak"] = line["tflops_reference_count"] * 1e12 / peak
            print(json.dumps(line), flush=True)
    ctx.set_core_func("wl")


if __name__ == "__main__":
    main()
