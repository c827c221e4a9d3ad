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
    """dram__bytes_read.sum + dram__bytes_write.sum of one pp2_kernel launch, from the committed ncu captures
    (profiles/pp2_dram_traffic.json, keyed "<sources>x<targets on this GPU>")."""
    try:
        with open(os.path.join(ROOT, "profiles", "pp2_dram_traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{n}x{nloc}")
    except (OSError, ValueError):
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------
def cpu_engine(threads=None, calibrate=None):
    """The reference's own points_affect_points<float,double> compiled from its sources (oracle/_ref/libo3d_ref_fast*.so:
    its templates, its stock -O3, OpenMP). The reference builds with -march=native (CMakeLists.txt:60); the GPU box's CPU is
    unknown at build time, so an x86-64-v3 (AVX2) and an x86-64-v4 (AVX-512) build ship and the faster one this host can run
    is used (`calibrate(engine) -> seconds` decides). Threads: every core this process may run on - from the affinity
    mask, NOT OMP_NUM_THREADS, which torchrun forces to 1. Falls back to the C restatement ("port") only if no reference
    build is present. Returns (engine, kind, threads, build description)."""
    from oracle import oracle_py
    threads = threads or oracle_py.host_threads()
    cands = []
    for tag, march in (("v4", "x86-64-v4"), ("v3", "x86-64-v3")):
        if tag == "v4" and not oracle_py.host_has_avx512():
            continue
        try:
            cands.append((oracle_py.Reference(fast=tag), f"libo3d_ref_fast{'_v4' if tag == 'v4' else ''}.so (reference templates, -O3 -march={march} -fopenmp)"))
        except Exception:
            pass
    if not cands:
        eng = oracle_py.Restatement()
        eng.set_threads(threads)
        return eng, "port", threads, "C restatement oracle/biot_oracle.c (-O2 -fopenmp)"
    for eng, _ in cands:
        eng.set_threads(threads)
    if calibrate is not None and len(cands) > 1:
        cands.sort(key=lambda c: calibrate(c[0]))
    return cands[0][0], "reference", threads, cands[0][1]


def cpu_reference_rate(n, x, s, r, seconds, steps=1, warmup=0, return_results=False):
    """That CPU code on a bounded sample: `nt` evenly strided targets against all n sources, nt sized so one step lasts
    about `seconds`. Returns (info, seconds per step, (sample indices, u, ug) | None)."""
    from omega3d_b200 import workloads as W

    def run(eng, nt):
        sel = W.strided_subset(n, nt)
        tx = np.ascontiguousarray(x[:, sel]); tr = np.ascontiguousarray(r[sel])
        tu, tug = np.zeros((3, sel.size), np.float32), np.zeros((9, sel.size), np.float32)
        t0 = time.perf_counter()
        eng.pts_on_pts(x, r, s, tx, tr, tu, tug)
        return time.perf_counter() - t0, sel, tu, tug

    from oracle import oracle_py
    cores = oracle_py.host_threads()
    nt0 = min(n, 4 * cores)
    eng, kind, cores, build = cpu_engine(cores, calibrate=lambda e: min(run(e, nt0)[0], run(e, nt0)[0]))
    t_cal, *_ = run(eng, nt0)
    t_cal = min(t_cal, run(eng, nt0)[0])
    rate0 = n * nt0 / max(t_cal, 1e-6)
    nt = int(min(n, max(nt0, cores * round(rate0 * seconds / n / cores))))
    for _ in range(warmup):
        run(eng, nt)
    times = []
    for _ in range(steps):
        dt, sel, tu, tug = run(eng, nt)
        times.append(dt)
    dt = float(np.mean(times))
    info = {"value": n * nt / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{nt} evenly strided targets x all {n} sources per step (~{dt:.1f} s), points_affect_points<float,double> {build}, "
                      f"{cores} OpenMP threads (affinity mask; OMP_NUM_THREADS ignored)"}
    return info, dt, (sel, tu, tug) if return_results else None


def parity_stats(gu, gg, ru, rg):
    """GPU (gu, gg) against reference (ru, rg) on one target sample. `*_err`: the max-norm relative error of SURVEY.md 8d,
    max|a-b| / max|b|. `*_rel_*`: PER-COMPONENT relative errors |a-b| / |b| over the components with |b| > 1e-3 max|b|
    (what north_star's "max relative error" reads as for components that are not in the rounding noise of a cancelling sum)."""
    out = {}
    for name, a, b in (("vel", gu, ru), ("grad", gg, rg)):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        scale = float(np.max(np.abs(b)))
        out[f"{name}_err"] = float(np.max(np.abs(a - b)) / scale)
        m = np.abs(b) > 1e-3 * scale
        rel = np.abs(a - b)[m] / np.abs(b)[m]
        out[f"{name}_rel_p50"] = float(np.percentile(rel, 50))
        out[f"{name}_rel_p99"] = float(np.percentile(rel, 99))
        out[f"{name}_rel_max"] = float(np.max(rel))
        out[f"{name}_components"] = int(m.sum())
    return out


def rank_parity(n, x_h, s_h, r_h, lo, hi, u, ug, steps, world, rank, count=256):
    """Every rank checks `count` evenly strided targets OF ITS OWN SHARD against the reference's CPU code (all n sources),
    on its share of the host cores. u, ug hold `steps` identical accumulated passes."""
    import torch
    from oracle import oracle_py
    from omega3d_b200 import workloads as W
    threads = max(1, oracle_py.host_threads() // world)
    eng, kind, threads, _ = cpu_engine(threads)
    sel = lo + W.strided_subset(hi - lo, count)
    tx = np.ascontiguousarray(x_h[:, sel]); tr = np.ascontiguousarray(r_h[sel])
    ru, rg = np.zeros((3, sel.size), np.float32), np.zeros((9, sel.size), np.float32)
    eng.pts_on_pts(x_h, r_h, s_h, tx, tr, ru, rg)
    idx = torch.from_numpy(sel - lo).to(u.device)
    gu = (u[:, idx] / steps).cpu().numpy()
    gg = (ug[:, idx] / steps).cpu().numpy()
    st = parity_stats(gu, gg, ru, rg)
    st["targets_checked"] = int(sel.size)
    st["against"] = kind
    return st


def workload_config(n, world):
    """The `config` object of both arms (ours and --impl reference describe the same workload)."""
    nloc = min(n, -(-(-(-n // world)) // 512) * 512)      # rank 0's shard: whole 512-particle tiles (omega3d_b200.device.shard_bounds)
    return {"workload": f"synthetic uniform vortex-particle cloud N={n} (positions U[-.5,.5]^3, strengths U/N, radius 1.5 N^-1/3), "
                        f"every particle source and target, vel+grad blob-on-blob WL core (north_star: 4M on one B200; BASELINE configs[4] sweep via --n)",
            "n_particles": n, "targets_per_gpu": nloc, "parallelism": f"targets sharded x{world}, sources all-gathered (NCCL)" if world > 1 else "single GPU",
            "l2": "256 MiB buffer written between timed iterations (L2 flush)"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from omega3d_b200 import workloads as W
    n = args.n
    x, s, r = W.random_cloud(n)
    # every step is a bounded sample of the workload, sized so that the whole run (warm-up included) stays near two minutes
    seconds = min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup))
    info, dt, _ = cpu_reference_rate(n, x, s, r, seconds=seconds, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": info["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 kernel / f64 accumulate", "data": "synthetic",
            "config": workload_config(n, max(1, args.gpus)),
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

    eng = DeviceBiotSavart(local_rank)
    props = eng.ctx.device_props(0)
    peak_probe_tf, _ = eng.probe_fp32_peak()
    eng.set_profiling(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    peaks = measured_peaks()
    f_max = float(peaks.get("sm_max_mhz", props["clock_khz"] / 1e3)) * 1e6
    peak_nominal = props["sm_count"] * 128 * 2 * f_max * 1e-12               # TFLOP/s, FP32 FMA at max SM clock

    def measure(n, steps, warmup, want_cpu, want_e2e):
        """One workload: `warmup` + `steps` device-resident steps (timed with CUDA events, max over ranks), the parity leg on
        every rank, optionally the CPU baseline (rank 0, one GPU) and the end-to-end leg through the host-pointer C ABI."""
        x_h, s_h, r_h = W.random_cloud(n)                     # identical on every rank (seeded)
        shard = ShardedBiotSavart(n, rank, world, eng)
        lo, hi = shard.lo, shard.hi
        nloc = hi - lo
        x = torch.from_numpy(np.ascontiguousarray(x_h[:, lo:hi])).to(dev)
        s = torch.from_numpy(np.ascontiguousarray(s_h[:, lo:hi])).to(dev)
        r = torch.from_numpy(np.ascontiguousarray(r_h[lo:hi])).to(dev)
        u = torch.zeros((3, nloc), dtype=torch.float32, device=dev)
        ug = torch.zeros((9, nloc), dtype=torch.float32, device=dev)
        shard.buffers(dev)                                    # allocations happen before anything is timed

        for _ in range(warmup):
            shard.step(x, s, r, u, ug)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        u.zero_(); ug.zero_()
        eng.launches = 0
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kernel_ms = []
        barrier()
        t_wall0 = time.perf_counter()
        for k in range(steps):
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
        value = float(n) * float(n) * steps / (total_ms * 1e-3)
        kern_ms_max = max_over_ranks(float(np.mean(kernel_ms)))

        # ---- parity of the device-resident result against the reference's CPU code: every rank, its own targets ----
        cpu_info, parity = None, None
        if not args.no_cpu:
            if world == 1 and want_cpu:
                # one GPU: the CPU-baseline sample (thousands of targets) doubles as the parity set
                cpu_info, _, (sel, ru, rg) = cpu_reference_rate(n, x_h, s_h, r_h, seconds=args.cpu_seconds, return_results=True)
                idx = torch.from_numpy(sel).to(dev)
                mine = parity_stats((u[:, idx] / steps).cpu().numpy(), (ug[:, idx] / steps).cpu().numpy(), ru, rg)
                mine["targets_checked"] = int(sel.size)
                mine["against"] = cpu_info["kind"]
            else:
                mine = rank_parity(n, x_h, s_h, r_h, lo, hi, u, ug, steps, world, rank)
            allp = [mine]
            if world > 1:
                allp = [None] * world
                dist.all_gather_object(allp, mine)
            parity = {k: max(p[k] for p in allp) for k in mine if k.endswith(("_err", "_p50", "_p99", "_max"))}
            parity.update({"targets_checked": sum(p["targets_checked"] for p in allp), "ranks": world, "against": mine["against"],
                           "vel_tol": 1e-5, "grad_tol": 1e-4, "norm": "*_err: max|a-b|/max|b| (the gate); *_rel_*: per-component |a-b|/|b| "
                           "over components with |b| > 1e-3 max|b|; max over ranks",
                           "ok": all(p["vel_err"] <= 1e-5 and p["grad_err"] <= 1e-4 for p in allp)})

        # ---- end to end through the host-pointer C ABI (the call the reference's gpu_cuda arm makes) ----
        e2e = None
        if want_e2e:
            def pinned(a):
                t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                return t, t.numpy()
            keep = [pinned(a) for a in (x_h, s_h, r_h, x_h[:, lo:hi], r_h[lo:hi], np.zeros((3, nloc), np.float32), np.zeros((9, nloc), np.float32))]
            hx, hs, hr, htx, htr, hu, hg = [k[1] for k in keep]
            long_steps = total_ms / steps > 5000.0              # (4 M on one GPU is 18.7 s a call: no separate warm-up call,
            e2e_steps = max(1, min(steps, 1 if long_steps else args.e2e_steps))   # its allocations are ~ms of the first one)
            if not (long_steps or args.e2e_no_warmup):
                eng.ctx.pts_on_pts(hx, hr, hs, htx, htr, hu, hg)     # warm-up (allocations)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                eng.ctx.pts_on_pts(hx, hr, hs, htx, htr, hu, hg)
            t_e2e = max_over_ranks(time.perf_counter() - t0)
            tm = eng.ctx.last_timing()
            e2e = {"value": float(n) * float(n) * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": (7 * n + 4 * nloc + 12 * nloc) * 4,
                   "d2h_bytes_per_step": 12 * nloc * 4, "steps": e2e_steps, "ms_per_step": t_e2e / e2e_steps * 1e3,
                   "kernel_ms": tm["kernel_ms"], "h2d_ms": tm["h2d_ms"], "d2h_ms": tm["d2h_ms"], "launches_per_step": tm["launches"],
                   "api": "o3d_cuda_pts_on_pts (include/o3d_cuda.h) with pinned host buffers"}
            # ... and with PAGEABLE host arrays - what the patched reference passes (std::vector storage): the context's
            # pinned staging ring carries them
            pg = [np.array(a, copy=True) for a in (x_h, s_h, r_h, x_h[:, lo:hi], r_h[lo:hi])] + [np.zeros((3, nloc), np.float32), np.zeros((9, nloc), np.float32)]
            pg_steps = 1 if long_steps else e2e_steps
            barrier()
            t0 = time.perf_counter()
            for _ in range(pg_steps):
                eng.ctx.pts_on_pts(pg[0], pg[2], pg[1], pg[3], pg[4], pg[5], pg[6])
            t_pg = max_over_ranks(time.perf_counter() - t0)
            tmp = eng.ctx.last_timing()
            e2e["pageable"] = {"value": float(n) * float(n) * pg_steps / t_pg, "steps": pg_steps, "ms_per_step": t_pg / pg_steps * 1e3,
                               "kernel_ms": tmp["kernel_ms"], "h2d_ms": tmp["h2d_ms"], "d2h_ms": tmp["d2h_ms"],
                               "of_pinned": (float(n) * float(n) * pg_steps / t_pg) / e2e["value"]}
        flops = float(n) * nloc * FLOPS_PER_INTERACTION + 12.0 * nloc
        achieved = flops / (kern_ms_max * 1e-3) * 1e-12
        return dict(n=n, nloc=nloc, value=value, total_ms=total_ms, steps=steps, launches=launches, clocks=clocks, t_wall=t_wall,
                    kern_ms=kern_ms_max, flops=flops, achieved=achieved, cpu_info=cpu_info, parity=parity, e2e=e2e)

    m = measure(args.n, args.steps, args.warmup, want_cpu=True, want_e2e=True)
    # the north star's largest point - 16 M particles on 8 GPUs - as ONE extra step after the headline workload
    big = None
    if world >= 8 and args.extra_n > args.n:
        big = measure(args.extra_n, 1, 0, want_cpu=False, want_e2e=False)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    n, nloc = m["n"], m["nloc"]
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m["total_ms"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 kernel / f64 accumulate", "data": "synthetic",
        "config": workload_config(n, world),
        "tflops_at_70": m["value"] * FLOPS_PER_INTERACTION * 1e-12,
        "roofline": {"bound": "fp32", "achieved": m["achieved"], "peak": peak_nominal, "unit": "TFLOP/s", "frac": m["achieved"] / peak_nominal,
                     "traffic": dram_traffic(n, nloc),
                     "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel at this shape "
                                     f"(profiles/pp2_dram_traffic.json; null = shape not captured); algorithmic bytes {32 * n + 16 * nloc + 96 * nloc}",
                     "peak_source": f"{props['sm_count']} SMs x 128 FP32 lanes x 2 x {f_max / 1e6:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                                    "MEASURED_PEAKS.json carries no FP32 figure - the path is FP32-pipe bound (arithmetic intensity ~1e6 flop/B), "
                                    "not HBM or tensor bound",
                     "peak_probe": peak_probe_tf, "frac_of_probe": m["achieved"] / peak_probe_tf if peak_probe_tf else None,
                     "probe": "packed-FMA (fma.rn.f32x2) issue loop timed on this GPU in this run",
                     "kernel": "o3d::pp2_kernel<2,true,384> (one persistent CTA per SM, static stream-K partition)", "kernel_ms": m["kern_ms"],
                     "flops_per_launch": m["flops"]},
        "e2e": m["e2e"],
        "gpu_launches": m["launches"], "gpu_launches_e2e_per_step": m["e2e"]["launches_per_step"],
        "clocks": m["clocks"], "wall_s_timed_region": m["t_wall"],
    }
    if m["cpu_info"] is not None:
        line["cpu_baseline"] = m["cpu_info"]
    if m["parity"] is not None:
        line["parity"] = m["parity"]
    if big is not None:
        line["extra"] = {"n16m": {"n_particles": big["n"], "targets_per_gpu": big["nloc"], "steps": 1, "value": big["value"], "unit": UNIT,
                                  "ms_per_step": big["total_ms"], "kernel_ms": big["kern_ms"], "frac": big["achieved"] / peak_nominal,
                                  "parity": big["parity"], "clocks": big["clocks"],
                                  "note": "one untimed-warm-up-free step of the 16 M north-star point after the headline workload; same kernels, same sharding"}}
    print(json.dumps(line), flush=True)


def run_inproc(args):
    """--inproc N: ONE process driving N GPUs through o3d_cuda_create(ndev = N) - what the single-process Omega3D does with
    O3D_CUDA_NDEV - with PAGEABLE host arrays, end to end through o3d_cuda_pts_on_pts: sources cross PCIe once (device 0) and
    reach the other devices as packed records over NVLink; every device uploads its target slice, computes, returns its
    slice. Prints one JSON line whose value IS the end-to-end rate (there is no device-resident leg in this mode)."""
    import torch
    from omega3d_b200 import influence as I
    from omega3d_b200 import workloads as W
    if not torch.cuda.is_available() or torch.cuda.device_count() < args.inproc:
        raise SystemExit(f"bench.py --inproc {args.inproc}: needs {args.inproc} CUDA devices")
    n, g = args.n, args.inproc
    x, s, r = W.random_cloud(n)
    u, ug = np.zeros((3, n), np.float32), np.zeros((9, n), np.float32)
    ctx = I.CudaContext(tuple(range(g)))
    samplers = [ClockSampler(k) for k in range(g)]
    for _ in range(max(1, min(args.warmup, 2))):
        ctx.pts_on_pts(x, r, s, x, r, u, ug)
    u[:] = 0; ug[:] = 0
    for sm in samplers:
        sm.start()
    t0 = time.perf_counter()
    ks = []
    for _ in range(args.steps):
        ctx.pts_on_pts(x, r, s, x, r, u, ug)
        ks.append(ctx.last_timing())
    dt = time.perf_counter() - t0
    clocks = [sm.stop() for sm in samplers]
    value = float(n) * float(n) * args.steps / dt
    parity = None
    if not args.no_cpu:
        sel = W.strided_subset(n, 512)
        eng, kind, threads, _ = cpu_engine()
        ru, rg = np.zeros((3, sel.size), np.float32), np.zeros((9, sel.size), np.float32)
        eng.pts_on_pts(x, r, s, np.ascontiguousarray(x[:, sel]), np.ascontiguousarray(r[sel]), ru, rg)
        parity = parity_stats(u[:, sel] / args.steps, ug[:, sel] / args.steps, ru, rg)
        parity.update({"targets_checked": int(sel.size), "against": kind, "vel_tol": 1e-5, "grad_tol": 1e-4,
                       "ok": parity["vel_err"] <= 1e-5 and parity["grad_err"] <= 1e-4})
    nloc = -(-n // g)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": g, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 kernel / f64 accumulate", "data": "synthetic", "impl": "ours-inproc",
            "config": {"workload": f"synthetic uniform vortex-particle cloud N={n}, vel+grad blob-on-blob WL core", "n_particles": n,
                       "parallelism": f"one process, one context over {g} GPUs (o3d_cuda_create ndev={g}); targets partitioned, sources "
                                      "uploaded once and replicated as packed records over NVLink", "host_memory": "pageable"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": (7 * n + 16 * n) * 4, "d2h_bytes_per_step": 12 * n * 4,
                    "kernel_ms": float(np.mean([k["kernel_ms"] for k in ks])), "h2d_ms": float(np.mean([k["h2d_ms"] for k in ks])),
                    "d2h_ms": float(np.mean([k["d2h_ms"] for k in ks])),
                    "api": "o3d_cuda_pts_on_pts (include/o3d_cuda.h), pageable host arrays, multi-device context"},
            "tflops_at_70": value * FLOPS_PER_INTERACTION * 1e-12,
            "gpu_launches": int(sum(k["launches"] for k in ks)), "targets_per_gpu": nloc,
            "clocks": clocks[0], "clocks_all": clocks}
    if parity is not None:
        line["parity"] = parity
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", "--n", dest="n", type=int, default=1 << 22,
                    help="particles (sources = targets); default 4 M = the north-star single-GPU size")
    ap.add_argument("--extra-n", type=int, default=1 << 24, help="with 8 GPUs: one extra step at this size (the 16 M north-star point); 0 = off")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="size of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-no-warmup", action="store_true", help="time the first end-to-end call too (very large N)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--inproc", type=int, default=0, help="N > 0: one process, one context over N GPUs, pageable host arrays (end to end only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.inproc > 0:
        run_inproc(args)
        return
    if world != args.gpus and rank == 0 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})", file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
