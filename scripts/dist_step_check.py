"""torchrun -> one process per GPU: ShardedConvection (NCCL all-gather of the packed records, every rank moving its own
tile-aligned block) against the single-GPU resident step. Tile-aligned blocks make the gathered record stream the
single-GPU stream, so the comparison is bit for bit. Rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 scripts/dist_step_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omega3d_b200 import convection as C  # noqa: E402
from omega3d_b200 import influence as I  # noqa: E402
from omega3d_b200 import workloads as W  # noqa: E402
from omega3d_b200.device import DeviceBiotSavart, ShardedConvection  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    order, steps, dt, fs = 2, 2, 0.01, (0.1, 0.0, 0.0)
    x, s, r = W.random_cloud(n, seed=77)
    s = (s * np.float32(30.0)).astype(np.float32)
    eng = DeviceBiotSavart(local)
    sc = ShardedConvection(n, rank, world, eng, order=order)
    lo, hi = sc.lo, sc.hi
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    xs, ss, rs, es = t(x[:, lo:hi]), t(s[:, lo:hi]), t(r[lo:hi]), torch.ones(hi - lo, device=dev)
    u, ug = torch.zeros((3, hi - lo), device=dev), torch.zeros((9, hi - lo), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        sc.advect(dt, fs, xs, ss, rs, es, u, ug)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # gather every rank's block on rank 0 (equal-size padded buffers)
    per = max(sc.sh.rec_per_rank, 1)
    mine = torch.zeros((19, per), device=dev)
    for k, a in enumerate((xs, ss, u, ug)):
        pass
    rows = torch.cat([xs, ss, es[None, :], u, ug], dim=0)
    mine[:, : hi - lo] = rows
    allb = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, allb, dst=0)
    if rank == 0:
        from omega3d_b200.device import shard_bounds
        got = torch.cat([b[:, : shard_bounds(n, world, k)[1] - shard_bounds(n, world, k)[0]] for k, b in enumerate(allb)], dim=1).cpu().numpy()
        p = C.DeviceParticles(I.CudaContext((local,))).upload(x, s, r)
        p.advect(order, 0.0, dt, fs, steps)
        o = p.download()
        ref = np.concatenate([o["x"], o["s"], o["elong"][None, :], o["u"], o["ug"]], axis=0)
        print(json.dumps({"check": "ShardedConvection over NCCL == single-GPU resident step", "ranks": world, "particles": n,
                          "order": order, "steps": steps, "bit_identical": bool(np.array_equal(got, ref)), "values": int(got.size), "values_differing": int(np.sum(got != ref)),
                          "max_abs_diff": float(np.max(np.abs(got - ref))), "ms_per_step_max_over_ranks": float(ms.item()) / steps}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
