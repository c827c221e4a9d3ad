#!/usr/bin/env python
"""bench.py - Biot-Savart interactions/s (velocity + 9 gradients, blob on blob, WL core) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n PARTICLES] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one pass of the hot path over one synthetic particle cloud: every particle is a source and a
target (the vort->vort find_vels of src/Convection.h:132-171 in the reference): pack the SoA sources into
32-byte records, [N > 1: NCCL all-gather of the packed records], one points-on-points launch.
Prints ONE JSON line (rank 0). Fields are described in DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_PER_INTERACTION = 70  # flops_0v_0bg: 54 + flops_tv_grads 16 (src/Kernels.h:155, src/CoreFunc.h:275)
METRIC = "Biot-Savart interactions/sec (vel+grad)"
UNIT = "interactions/s"


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU every 100 ms through NVML while a region runs."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, torch_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.ok = [], 0, None, False
        self.power = []
        self._halt = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(np.min(self.samples)),
                "sm_max_mhz": self.max_mhz, "power_w_median": float(np.median(self.power)) if self.power else None,
                "samples": len(self.samples), "reasons": [n for b, n in self.BITS.items() if self.reasons & b]}


def dram_traffic(n, nloc):
    """dram__bytes_read.sum + dram__bytes_write.sum of one pp2_kernel launch, from the committed ncu --set full
    captures (profiles/pp2_dram_traffic.json, keyed by particle count; single-GPU launches only)."""
    try:
        with open(os.path.join(ROOT, "profiles", "pp2_dram_traffic.json")) as f:
            t = json.load(f)
        return t.get(str(n)) if n == nloc else None
    except (OSError, ValueError):
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n, x, s, r, seconds, steps=1, warmup=0, return_results=False):
    """The reference's own points_affect_points<float,double> (oracle/_ref/libo3d_ref_fast.so: its templates,
    stock -O3 flags, OpenMP on every host core) on a bounded sample: `nt` evenly strided targets against all n
    sources, nt sized so one step lasts about `seconds`. Falls back to the C restatement ("port") only if the
    reference build is absent. Returns (rate, info, sample indices, results)."""
    from oracle import oracle_py
    from omega3d_b200 import workloads as W
    try:
        eng, kind = oracle_py.Reference(fast=True), "reference"
    except Exception:
        eng, kind = oracle_py.Restatement(), "port"
    cores = eng.max_threads()

    def run(nt):
        sel = W.strided_subset(n, nt)
        tx = np.ascontiguousarray(x[:, sel]); tr = np.ascontiguousarray(r[sel])
        tu, tug = np.zeros((3, sel.size), np.float32), np.zeros((9, sel.size), np.float32)
        t0 = time.perf_counter()
        eng.pts_on_pts(x, r, s, tx, tr, tu, tug)
        return time.perf_counter() - t0, sel, tu, tug

    nt0 = min(n, 4 * cores)
    t_cal, *_ = run(nt0)
    t_cal, *_ = run(nt0)
    rate0 = n * nt0 / max(t_cal, 1e-6)
    nt = int(min(n, max(nt0, cores * round(rate0 * seconds / n / cores))))
    for _ in range(warmup):
        run(nt)
    times = []
    for _ in range(steps):
        dt, sel, tu, tug = run(nt)
        times.append(dt)
    dt = float(np.mean(times))
    info = {"value": n * nt / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{nt} evenly strided targets x all {n} sources per step (~{dt:.1f} s), points_affect_points<float,double> "
                      f"{'libo3d_ref_fast.so (reference templates, -O3 -march=x86-64-v3 -fopenmp)' if kind == 'reference' else 'C restatement'}"}
    return info, dt, (sel, tu, tug) if return_results else None


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    from omega3d_b200 import workloads as W
    n = args.n
    x, s, r = W.random_cloud(n)
    info, dt, _ = cpu_reference_rate(n, x, s, r, seconds=args.cpu_seconds, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": info["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 kernel / f64 accumulate", "data": "synthetic",
            "config": {"workload": f"synthetic uniform vortex-particle cloud N={n}, vel+grad blob-on-blob WL core (bounded target sample)",
                       "n_particles": n},
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from omega3d_b200 import workloads as W
    from omega3d_b200.device import DeviceBiotSavart, ShardedBiotSavart

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - omega3d_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    x_h, s_h, r_h = W.random_cloud(n)                     # identical on every rank (seeded)
    eng = DeviceBiotSavart(local_rank)
    shard = ShardedBiotSavart(n, rank, world, eng)
    lo, hi = shard.lo, shard.hi
    nloc = hi - lo
    x = torch.from_numpy(np.ascontiguousarray(x_h[:, lo:hi])).to(dev)
    s = torch.from_numpy(np.ascontiguousarray(s_h[:, lo:hi])).to(dev)
    r = torch.from_numpy(np.ascontiguousarray(r_h[lo:hi])).to(dev)
    u = torch.zeros((3, nloc), dtype=torch.float32, device=dev)
    ug = torch.zeros((9, nloc), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    props = eng.ctx.device_props(0)
    peak_probe_tf, _ = eng.probe_fp32_peak()
    eng.set_profiling(True)

    # ---- device-resident steps ----
    for _ in range(args.warmup):
        shard.step(x, s, r, u, ug)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    u.zero_(); ug.zero_()
    eng.launches = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                      # L2 flush between timed iterations (outside the events)
        ev[k][0].record()
        shard.step(x, s, r, u, ug)
        ev[k][1].record()
        kernel_ms.append(eng.last_kernel_ms())             # waits for this step's dominant kernel
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(sum(step_ms))
    launches = eng.launches
    clocks = sampler.stop()
    value = float(n) * float(n) * args.steps / (total_ms * 1e-3)
    kern_ms = float(np.mean(kernel_ms))
    kern_ms_max = max_over_ranks(kern_ms)

    # ---- parity of the device-resident result against the CPU reference sample (rank 0, N=1 only) ----
    cpu_info, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_info, _, (sel, ru, rg) = cpu_reference_rate(n, x_h, s_h, r_h, seconds=args.cpu_seconds, return_results=True)
        gu = (u[:, torch.from_numpy(sel).to(dev)] / args.steps).cpu().numpy()   # u accumulated `steps` identical passes
        gg = (ug[:, torch.from_numpy(sel).to(dev)] / args.steps).cpu().numpy()
        parity = {"targets_checked": int(sel.size), "vel_err": float(np.max(np.abs(gu - ru)) / np.max(np.abs(ru))),
                  "grad_err": float(np.max(np.abs(gg - rg)) / np.max(np.abs(rg))), "vel_tol": 1e-5, "grad_tol": 1e-4,
                  "against": cpu_info["kind"]}

    # ---- end to end through the host-pointer C ABI (the call the reference's gpu_cuda arm makes) ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = [pinned(a) for a in (x_h, s_h, r_h, x_h[:, lo:hi], r_h[lo:hi], np.zeros((3, nloc), np.float32), np.zeros((9, nloc), np.float32))]
    hx, hs, hr, htx, htr, hu, hg = [k[1] for k in keep]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if not args.e2e_no_warmup:                           # (the 16M sweep point skips it: one call there is ~40 s)
        eng.ctx.pts_on_pts(hx, hr, hs, htx, htr, hu, hg)     # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.ctx.pts_on_pts(hx, hr, hs, htx, htr, hu, hg)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    tm = eng.ctx.last_timing()
    e2e_launches = tm["launches"]
    h2d = (7 * n + 4 * nloc + 12 * nloc) * 4
    d2h = 12 * nloc * 4
    e2e_value = float(n) * float(n) * e2e_steps / t_e2e

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    peaks = measured_peaks()
    f_max = float(peaks.get("sm_max_mhz", props["clock_khz"] / 1e3)) * 1e6
    peak_nominal = props["sm_count"] * 128 * 2 * f_max * 1e-12               # TFLOP/s, FP32 FMA at max SM clock
    achieved = (float(n) * nloc * FLOPS_PER_INTERACTION + 12.0 * nloc) / (kern_ms_max * 1e-3) * 1e-12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 kernel / f64 accumulate", "data": "synthetic",
        "config": {"workload": f"synthetic uniform vortex-particle cloud N={n} (positions U[-.5,.5]^3, strengths U/N, radius 1.5 N^-1/3), "
                               f"every particle source and target, vel+grad blob-on-blob WL core (BASELINE configs[1] size; configs[4] sweep via --n)",
                   "n_particles": n, "targets_per_gpu": nloc, "parallelism": f"targets sharded x{world}, sources all-gathered (NCCL)" if world > 1 else "single GPU",
                   "l2": "256 MiB buffer written between timed iterations (L2 flush)"},
        "tflops_at_70": value * FLOPS_PER_INTERACTION * 1e-12,
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s", "frac": achieved / peak_nominal,
                     "traffic": dram_traffic(n, nloc),
                     "traffic_note": "ncu dram bytes read+write of one launch (profiles/pp2_dram_traffic.json); algorithmic bytes "
                                     f"{32 * n + 16 * nloc + 96 * nloc}; at 1M the launch splits the sources 4 ways to fill its last wave and "
                                     "writes 403 MB of FP64 partial slabs (0.06 ms at HBM speed in a 1170 ms launch)",
                     "peak_source": f"{props['sm_count']} SMs x 128 FP32 lanes x 2 x {f_max / 1e6:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                                    "MEASURED_PEAKS.json carries no FP32 figure - the path is FP32-pipe bound (arithmetic intensity ~1e6 flop/B), "
                                    "not HBM or tensor bound",
                     "peak_probe": peak_probe_tf, "frac_of_probe": achieved / peak_probe_tf if peak_probe_tf else None,
                     "probe": "packed-FMA (fma.rn.f32x2) issue loop timed on this GPU in this run",
                     "kernel": "o3d::pp2_kernel<2,true,128>", "kernel_ms": kern_ms_max,
                     "flops_per_launch": float(n) * nloc * FLOPS_PER_INTERACTION + 12.0 * nloc},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": t_e2e / e2e_steps * 1e3, "kernel_ms": tm["kernel_ms"], "h2d_ms": tm["h2d_ms"], "d2h_ms": tm["d2h_ms"],
                "api": "o3d_cuda_pts_on_pts (include/o3d_cuda.h) with pinned host buffers"},
        "gpu_launches": launches, "gpu_launches_e2e_per_step": e2e_launches,
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if cpu_info is not None:
        line["cpu_baseline"] = cpu_info
        line["parity"] = parity
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", "--n", dest="n", type=int, default=1 << 20, help="particles (sources = targets)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="size of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-no-warmup", action="store_true", help="time the first end-to-end call too (very large N)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and rank == 0 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})", file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
