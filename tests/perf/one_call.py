"""One evaluation (vel+grad) on a uniform cloud - the thing to put under ncu when only the main kernel's counters at a given
size are wanted:  python tests/perf/one_call.py N [targets] [dev]
default: o3d_cuda_pts_on_pts with host arrays (the kernel stores FP64 sums, pp_accumulate_kernel adds them to the uploaded outputs);
dev: the device-resident step bench.py times (omega3d_b200.device.ShardedBiotSavart: outputs read-modify-written by the kernel)."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from omega3d_b200 import influence as I   # noqa: E402
from omega3d_b200 import workloads as W   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
nt = int(sys.argv[2]) if len(sys.argv) > 2 else n
x, s, r = W.random_cloud(n)
if "dev" in sys.argv[3:]:
    import torch
    from omega3d_b200.device import DeviceBiotSavart, ShardedBiotSavart
    world = max(1, n // nt)                      # rank 0 of `world` ranks holds the first nt particles as its targets
    eng = DeviceBiotSavart(0)
    eng.set_profiling(True)
    sh = ShardedBiotSavart(n, 0, 1, eng)
    dev = torch.device("cuda", 0)
    xs, ss, rs = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, s, r))
    ud, ugd = torch.zeros((3, nt), device=dev), torch.zeros((9, nt), device=dev)
    sh.buffers(dev)
    if nt == n:
        sh.step(xs, ss, rs, ud, ugd)
    else:                                        # a shard's launch: all sources, the first nt particles as targets
        packed = eng.pack(xs, ss, rs)
        eng.pts_on_pts(packed, xs[:, :nt].contiguous(), rs[:nt].contiguous(), ud, ugd)
    torch.cuda.synchronize()
    print(f"N={n} targets={nt} device-resident kernel {eng.last_kernel_ms():.3f} ms  |u|max {float(ud.abs().max()):.4e}")
    sys.exit(0)
u, ug = np.zeros((3, nt), np.float32), np.zeros((9, nt), np.float32)
ctx = I.CudaContext((0,))
ctx.pts_on_pts(x, r, s, np.ascontiguousarray(x[:, :nt]), np.ascontiguousarray(r[:nt]), u, ug)
t = ctx.last_timing()
print(f"N={n} targets={nt} kernel {t['kernel_ms']:.3f} ms  launches {t['launches']}  |u|max {np.abs(u).max():.4e}")
