"""Device-resident evaluation: particle arrays already in HBM, one process per GPU.

PyTorch is plumbing here - device memory, the current CUDA stream and ``torch.distributed`` (NCCL over
NVLink) - and nothing else: every kernel launched is this repo's own, through the ``*_dev`` entry points of
``include/o3d_cuda.h``. The multi-GPU scheme is SURVEY.md section 8e: targets are block-partitioned across
ranks, each rank packs the sources it owns into 32-byte records, one all-gather replicates the packed
records, and every rank evaluates its own targets against all records. No reduction is needed.
"""
from __future__ import annotations

from ctypes import byref, c_double, c_void_p

import torch

from . import _lib
from .influence import CudaContext

REC_FLOATS = 8  # one packed source record = 2 x float4


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


class DeviceBiotSavart:
    """particles -> points on one GPU with torch tensors as the SoA containers.

    Tensors are float32, C-contiguous rows: x (3,n), s (3,n), r (n,), u (3,n), ug (9,n)."""

    def __init__(self, device: int = 0, ctx: CudaContext | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("omega3d_b200.device needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = ctx or CudaContext((device,))
        self.lib = self.ctx.lib
        self.launches = 0

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def packed_records(self, ns: int) -> int:
        return int(self.lib.o3d_cuda_packed_records(ns))

    def pack(self, x, s, r, out=None):
        """SoA sources -> packed record stream; the whole of `out` is filled (zero-strength records past ns)."""
        ns = x.shape[1]
        if out is None:
            out = torch.empty(self.packed_records(ns) * REC_FLOATS, dtype=torch.float32, device=self.device)
        nrec = out.numel() // REC_FLOATS
        assert x.is_contiguous() and s.is_contiguous()
        self.ctx.check(self.lib.o3d_cuda_pack_sources_dev(self.ctx.h, self._stream(), ns, _p(x[0]), _p(x[1]), _p(x[2]), _p(r),
                                                          _p(s[0]), _p(s[1]), _p(s[2]), nrec, _p(out)))
        self.launches += 1
        return out

    def pts_on_pts(self, packed, tx, tr, u, ug):
        """u (3,nt) += , ug (9,nt) += (None: velocity only); tr None: singular targets."""
        nrec = packed.numel() // REC_FLOATS
        nt = tx.shape[1]
        self.ctx.check(self.lib.o3d_cuda_pts_on_pts_dev(self.ctx.h, self._stream(), nrec, _p(packed), nt, _p(tx[0]), _p(tx[1]),
                                                        _p(tx[2]), _p(tr), _p(u[0]), _p(u[1]), _p(u[2]), _p(ug),
                                                        ug.stride(0) if ug is not None else 0))
        self.launches += self.ctx.last_timing()["launches"]

    def set_profiling(self, on: bool):
        self.ctx.check(self.lib.o3d_cuda_set_profiling(self.ctx.h, int(on)))

    def last_kernel_ms(self) -> float:
        ms = c_double()
        self.ctx.check(self.lib.o3d_cuda_dev_kernel_ms(self.ctx.h, byref(ms)))
        return ms.value

    def probe_fp32_peak(self):
        tf, ms = c_double(), c_double()
        self.ctx.check(self.lib.o3d_cuda_probe_fp32_peak(self.ctx.h, byref(tf), byref(ms)))
        return tf.value, ms.value


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block partition, the same rule as the C ABI's in-process partition (capi.cu: partition)."""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))


class ShardedBiotSavart:
    """One rank of a target-sharded evaluation. Each rank holds the particles [lo, hi) as BOTH its share
    of the sources and its targets; ``step`` = pack local sources -> all-gather packed records -> evaluate.

    ``backend`` collectives go through ``torch.distributed`` (NCCL on GPUs; the host-side bookkeeping of this
    class is exercised on CPU with gloo in tests/test_sharding.py)."""

    def __init__(self, n_total: int, rank: int, world: int, engine: DeviceBiotSavart | None):
        self.n, self.rank, self.world, self.engine = n_total, rank, world, engine
        self.lo, self.hi = shard_bounds(n_total, world, rank)
        per = (n_total + world - 1) // world
        # every rank contributes the same number of records so one all_gather_into_tensor suffices;
        # short ranks pad with zero-strength records (exactly zero contribution)
        self.rec_per_rank = int(_lib.load().o3d_cuda_packed_records(per))
        self.local_packed = None
        self.gathered = None

    def buffers(self, device):
        if self.local_packed is None:
            self.local_packed = torch.empty(self.rec_per_rank * REC_FLOATS, dtype=torch.float32, device=device)
            self.gathered = (self.local_packed if self.world == 1 else
                             torch.empty(self.world * self.rec_per_rank * REC_FLOATS, dtype=torch.float32, device=device))
        return self.local_packed, self.gathered

    def step(self, x, s, r, u, ug):
        """x,s (3,nloc), r (nloc,): this rank's particles; u (3,nloc), ug (9,nloc) accumulate."""
        import torch.distributed as dist
        eng = self.engine
        local, gathered = self.buffers(x.device)
        eng.pack(x, s, r, out=local)   # fills all rec_per_rank records; those past this rank's count have zero strength
        if self.world > 1:
            dist.all_gather_into_tensor(gathered, local)
        eng.pts_on_pts(gathered, x, r, u, ug)
